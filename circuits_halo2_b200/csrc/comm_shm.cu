// A ready-made `sb_comm` for the ranks of ONE box (one process per GPU), with no collective library and no interpreter on the data path:
//   * small host records (the 128-byte partial commitments of a window-sharded MSM, IPC handles) meet in a POSIX shared-memory mailbox: every rank
//     posts into its slot, bumps its sequence number and spins until the others have posted the same generation -- microseconds per exchange instead
//     of the hundreds of a Python callback into torch.distributed;
//   * device buffers move by direct peer copies over NVLink / NVSwitch through CUDA IPC mappings of the peers' buffers (pull model: a rank copies the
//     parts it needs out of its peers' memory on its own stream), bracketed by mailbox barriers.
// The library still links nothing but cudart.  NCCL (plonk.py::ShardComm) remains available through the generic callback interface; this
// implementation is what bench.py uses for sharded proofs.  North-star wording: "partial G1 sums are reduced on the host", exchanges over NVLink.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <map>
#include <string>
#include <thread>

#include "common.cuh"

namespace {

const int MAX_WORLD = 8;
const size_t SLOT_BYTES = 16 << 10;

struct Mailbox {
    std::atomic<uint64_t> seq[MAX_WORLD];          // generation each rank has posted
    std::atomic<uint32_t> attached;                // ranks that have mapped the segment (destroy: the last one unlinks)
    uint8_t pad[64 - sizeof(std::atomic<uint32_t>)];
    uint8_t data[2][MAX_WORLD][SLOT_BYTES];        // two generations in flight at most (see allgather)
};

struct IpcKey {
    uint8_t h[sizeof(cudaIpcMemHandle_t)];
    bool operator<(const IpcKey &o) const { return memcmp(h, o.h, sizeof h) < 0; }
};

}  // namespace

struct sb_shm_comm {
    sb_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    std::string name;
    Mailbox *mb = nullptr;
    uint64_t gen = 0;
    std::map<IpcKey, void *> mapped[MAX_WORLD];  // peer allocations opened so far
    double timeout_s = 120.0;
};

namespace {

int32_t host_allgather(sb_shm_comm *c, const void *send, void *recv, size_t bytes) {
    if (bytes > SLOT_BYTES) { sb::set_last_error("sb_comm_shm: host record of %zu bytes exceeds the %zu-byte mailbox slot", bytes, SLOT_BYTES); return 1; }
    Mailbox *mb = c->mb;
    const uint64_t g = ++c->gen;
    const int par = (int)(g & 1);
    if (bytes) memcpy(mb->data[par][c->rank], send, bytes);
    mb->seq[c->rank].store(g, std::memory_order_release);
    // A rank can be at most one generation ahead of the slowest one (it cannot finish generation g + 1 before everybody has posted g + 1, which
    // they do only after reading generation g), so two data buffers are enough.
    const auto t0 = std::chrono::steady_clock::now();
    for (int q = 0; q < c->world; q++) {
        uint32_t spins = 0;
        while (mb->seq[q].load(std::memory_order_acquire) < g) {
            if (++spins > 2000) {
                std::this_thread::yield();
                if ((spins & 0xfff) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > c->timeout_s) {
                    sb::set_last_error("sb_comm_shm: rank %d waited %.0f s for rank %d (generation %llu)", c->rank, c->timeout_s, q, (unsigned long long)g);
                    return 1;
                }
            }
        }
        if (bytes && recv) memcpy((uint8_t *)recv + (size_t)q * bytes, mb->data[par][q], bytes);
    }
    return 0;
}

int32_t cb_allgather_host(void *user, const void *send, void *recv, size_t bytes) { return host_allgather((sb_shm_comm *)user, send, recv, bytes); }

// every rank's mapping of every rank's buffer `d_local` (same call on all ranks): peers[q] addresses rank q's buffer in THIS process
int32_t map_peers(sb_shm_comm *c, void *d_local, void *peers[MAX_WORLD]) {
    struct Rec { cudaIpcMemHandle_t h; uint64_t off; } mine, all[MAX_WORLD];
    memset(&mine, 0, sizeof mine);
    // the handle names the whole allocation: find its base through a pointer-attribute query of the range
    void *base = nullptr;
    size_t range = 0;
    {
        typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
        static range_fn get_range = nullptr;
        static bool tried = false;
        if (!tried) {
            tried = true;
            if (void *drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL)) get_range = (range_fn)dlsym(drv, "cuMemGetAddressRange_v2");
        }
        unsigned long long b = 0;
        if (get_range && get_range(&b, &range, (unsigned long long)(uintptr_t)d_local) == 0) base = (void *)(uintptr_t)b;
        else base = d_local;  // our callers pass allocation bases (scratch slots)
    }
    if (cudaIpcGetMemHandle(&mine.h, base) != cudaSuccess) { sb::set_last_error("sb_comm_shm: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    mine.off = (uint64_t)((uint8_t *)d_local - (uint8_t *)base);
    if (host_allgather(c, &mine, all, sizeof mine) != 0) return 1;
    for (int q = 0; q < c->world; q++) {
        if (q == c->rank) { peers[q] = d_local; continue; }
        IpcKey key;
        memcpy(key.h, &all[q].h, sizeof key.h);
        auto it = c->mapped[q].find(key);
        void *p = nullptr;
        if (it != c->mapped[q].end()) {
            p = it->second;
        } else {
            if (cudaIpcOpenMemHandle(&p, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                sb::set_last_error("sb_comm_shm: cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
            c->mapped[q][key] = p;
        }
        peers[q] = (uint8_t *)p + all[q].off;
    }
    return 0;
}

// in place: rank r's part already sits at d_buf + r * bytes; pull everybody else's part out of their buffers
int32_t cb_allgather_dev(void *user, void *d_buf, size_t bytes, void *stream) {
    sb_shm_comm *c = (sb_shm_comm *)user;
    cudaStream_t st = (cudaStream_t)stream;
    void *peers[MAX_WORLD];
    if (map_peers(c, d_buf, peers) != 0) return 1;   // also a barrier: every rank's part is complete (callers synchronise their stream first)
    for (int i = 1; i < c->world; i++) {
        const int q = (c->rank + i) % c->world;      // staggered: not everybody reads from rank 0 first
        if (cudaMemcpyAsync((uint8_t *)d_buf + (size_t)q * bytes, (const uint8_t *)peers[q] + (size_t)q * bytes, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            sb::set_last_error("sb_comm_shm: peer copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { sb::set_last_error("sb_comm_shm: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    return host_allgather(c, nullptr, nullptr, 0);   // nobody overwrites its part while a peer is still reading it
}

// block q of d_send goes to rank q; block q of d_recv comes from rank q
int32_t cb_alltoall_dev(void *user, const void *d_send, void *d_recv, size_t bytes, void *stream) {
    sb_shm_comm *c = (sb_shm_comm *)user;
    cudaStream_t st = (cudaStream_t)stream;
    void *peers[MAX_WORLD];
    if (map_peers(c, const_cast<void *>(d_send), peers) != 0) return 1;
    for (int i = 0; i < c->world; i++) {
        const int q = (c->rank + i) % c->world;
        if (cudaMemcpyAsync((uint8_t *)d_recv + (size_t)q * bytes, (const uint8_t *)peers[q] + (size_t)c->rank * bytes, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            sb::set_last_error("sb_comm_shm: peer copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { sb::set_last_error("sb_comm_shm: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    return host_allgather(c, nullptr, nullptr, 0);
}

}  // namespace

extern "C" {

int32_t sb_comm_shm_create(sb_ctx *ctx, const char *name, int32_t rank, int32_t world, sb_comm *out_comm, sb_shm_comm **out_handle) {
    if (!name || !out_comm || !out_handle) return SB_ERR_ARG;
    SB_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "sb_comm_shm_create: 1 <= world <= 8, 0 <= rank < world");
    sb_shm_comm *c = new sb_shm_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    c->name = std::string("/sb_b200_") + name;
    const int fd = shm_open(c->name.c_str(), O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, sizeof(Mailbox)) != 0) {
        sb::set_last_error("sb_comm_shm_create: shm_open / ftruncate(%s) failed", c->name.c_str());
        if (fd >= 0) close(fd);
        delete c;
        return SB_ERR_ALLOC;
    }
    void *p = mmap(nullptr, sizeof(Mailbox), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        sb::set_last_error("sb_comm_shm_create: mmap failed");
        delete c;
        return SB_ERR_ALLOC;
    }
    c->mb = (Mailbox *)p;   // a fresh segment is zero-filled: sequence numbers start at 0 on every rank
    c->mb->attached.fetch_add(1);
    if (ctx) sb::ctx_retain(ctx);
    out_comm->rank = rank;
    out_comm->world = world;
    out_comm->user = c;
    out_comm->allgather_host = cb_allgather_host;
    out_comm->allgather_dev = cb_allgather_dev;
    out_comm->alltoall_dev = cb_alltoall_dev;
    // first exchange: everybody is attached before anybody proceeds (and before anybody could unlink the name)
    if (host_allgather(c, nullptr, nullptr, 0) != 0) {
        munmap(c->mb, sizeof(Mailbox));
        if (ctx) sb::ctx_release(ctx);
        delete c;
        return SB_ERR_ARG;
    }
    *out_handle = c;
    return SB_OK;
}

int32_t sb_comm_shm_destroy(sb_shm_comm *c) {
    if (!c) return SB_OK;
    for (int q = 0; q < MAX_WORLD; q++)
        for (auto &kv : c->mapped[q]) cudaIpcCloseMemHandle(kv.second);
    if (c->mb) {
        if (c->mb->attached.fetch_sub(1) == 1) shm_unlink(c->name.c_str());
        munmap(c->mb, sizeof(Mailbox));
    }
    if (c->ctx) sb::ctx_release(c->ctx);
    delete c;
    return SB_OK;
}

}  // extern "C"
