// nt_hostframe.cpp — one RGBA8 host frame shared by the ranks of a one-process-per-GPU render (include/nettracer_b200.h
// nt_host_frame_*).  A POSIX shared-memory segment, mapped by every rank and page-locked for its GPU
// (cudaHostRegister, portable + mapped), so that nt_render's zero-copy path stores each rank's row bands straight into
// the frame rank 0's host reads.  The first page holds one flag line per rank ("my shard of frame seq is complete") and
// rank 0's acknowledgement ("frame seq has been consumed"); they are plain host atomics - the GPU side of every rank
// has already been synchronised by the blocking nt_render when a flag is posted.
// Replaces the reference's network tile gather, which cannot be cited (/root/reference/README:1-3 holds no code).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>

#include "../../include/nettracer_b200.h"

int nt_fail_public(int code, const char *fmt, ...); // nt_api.cu

namespace {
constexpr uint32_t kMagic = 0x4e544846u; // "NTHF"
constexpr size_t kHeaderBytes = 8192;    // flag lines (64 bytes each): [0] header, [1] ack, [2 + r] rank r; up to 126 ranks
constexpr uint32_t kMaxRanks = 126;
struct Line { std::atomic<uint32_t> v; uint32_t pad[15]; };
static_assert(sizeof(Line) == 64, "one cache line per flag");
static_assert(std::atomic<uint32_t>::is_always_lock_free, "flags must be plain words");
} // namespace

struct nt_host_frame {
    std::string name;
    uint8_t *base = nullptr;
    size_t map_bytes = 0, frame_bytes = 0;
    uint32_t n_ranks = 0;
    bool registered = false;
    Line *line(uint32_t i) const { return (Line *)base + i; }
};

extern "C" void nt_host_frame_close(nt_host_frame *f, int unlink_segment) {
    if (!f) return;
    if (f->registered) cudaHostUnregister(f->base);
    if (f->base) munmap(f->base, f->map_bytes);
    if (unlink_segment) shm_unlink(f->name.c_str());
    cudaGetLastError();
    delete f;
}

extern "C" int nt_host_frame_open(const char *name, size_t frame_bytes, uint32_t n_ranks, int create, int device, nt_host_frame **out) {
    if (!out) return nt_fail_public(NT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!name || name[0] != '/' || frame_bytes == 0 || n_ranks == 0 || n_ranks > kMaxRanks) return nt_fail_public(NT_ERR_INVALID, "bad argument (name must start with '/', 1..%u ranks)", kMaxRanks);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return nt_fail_public(NT_ERR_NO_DEVICE, "device %d is not usable", device); }
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t map_bytes = (kHeaderBytes + frame_bytes + page - 1) / page * page;
    int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0 && create && errno == EEXIST) { // a stale segment of a crashed run
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    }
    if (fd < 0) return nt_fail_public(NT_ERR_SYSTEM, "shm_open(%s): %s", name, strerror(errno));
    if (create && ftruncate(fd, (off_t)map_bytes) != 0) {
        const int e = errno;
        close(fd); shm_unlink(name);
        return nt_fail_public(NT_ERR_SYSTEM, "ftruncate(%s, %zu): %s", name, map_bytes, strerror(e));
    }
    if (!create) {
        struct stat sb;
        if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < map_bytes) { close(fd); return nt_fail_public(NT_ERR_INVALID, "segment %s is smaller than %zu bytes", name, map_bytes); }
    }
    void *p = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
    const int me = errno;
    close(fd);
    if (p == MAP_FAILED) {
        if (create) shm_unlink(name);
        return nt_fail_public(NT_ERR_SYSTEM, "mmap(%s): %s", name, strerror(me));
    }
    nt_host_frame *f = new (std::nothrow) nt_host_frame;
    if (!f) { munmap(p, map_bytes); return nt_fail_public(NT_ERR_NOMEM, "out of host memory"); }
    f->name = name; f->base = (uint8_t *)p; f->map_bytes = map_bytes; f->frame_bytes = frame_bytes; f->n_ranks = n_ranks;
    if (create) {
        memset(p, 0, kHeaderBytes);
        ((uint32_t *)p)[1] = n_ranks;
        std::atomic_thread_fence(std::memory_order_release);
        ((std::atomic<uint32_t> *)p)->store(kMagic, std::memory_order_release);
    } else if (((std::atomic<uint32_t> *)p)->load(std::memory_order_acquire) != kMagic || ((uint32_t *)p)[1] != n_ranks) {
        nt_host_frame_close(f, 0);
        return nt_fail_public(NT_ERR_INVALID, "segment %s was not made by nt_host_frame_open for %u ranks", name, n_ranks);
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaHostRegister(p, map_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        nt_host_frame_close(f, create);
        return nt_fail_public(NT_ERR_CUDA, "cudaHostRegister(%zu bytes): %s", map_bytes, cudaGetErrorString(e));
    }
    f->registered = true;
    *out = f;
    return NT_OK;
}

extern "C" uint8_t *nt_host_frame_pixels(nt_host_frame *f) { return f ? f->base + kHeaderBytes : nullptr; }
extern "C" uint32_t *nt_host_frame_flag(nt_host_frame *f, uint32_t rank) {
    return f && rank < f->n_ranks ? (uint32_t *)&f->line(2 + rank)->v : nullptr;
}

static int wait_line(Line *l, uint32_t seq, uint32_t timeout_ms) {
    if ((int32_t)(l->v.load(std::memory_order_acquire) - seq) >= 0) return NT_OK;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; ++spins) {
        if ((int32_t)(l->v.load(std::memory_order_acquire) - seq) >= 0) return NT_OK;
        if ((spins & 255u) == 255u) {
            const auto dt = std::chrono::steady_clock::now() - t0;
            if (dt > std::chrono::milliseconds(timeout_ms)) return NT_ERR_TIMEOUT;
            if (dt > std::chrono::milliseconds(2)) std::this_thread::yield(); // long waits (another rank builds a BVH) must not burn a core another rank needs
        }
    }
}

extern "C" int nt_host_frame_post(nt_host_frame *f, uint32_t rank, uint32_t seq) {
    if (!f || rank >= f->n_ranks) return nt_fail_public(NT_ERR_INVALID, "bad rank");
    f->line(2 + rank)->v.store(seq, std::memory_order_release);
    return NT_OK;
}

extern "C" int nt_host_frame_wait_all(nt_host_frame *f, uint32_t seq, uint32_t timeout_ms) {
    if (!f) return nt_fail_public(NT_ERR_INVALID, "NULL frame");
    for (uint32_t r = 0; r < f->n_ranks; ++r)
        if (wait_line(f->line(2 + r), seq, timeout_ms) != NT_OK) return nt_fail_public(NT_ERR_TIMEOUT, "rank %u has not posted frame %u within %u ms", r, seq, timeout_ms);
    return NT_OK;
}

extern "C" int nt_host_frame_ack(nt_host_frame *f, uint32_t seq) {
    if (!f) return nt_fail_public(NT_ERR_INVALID, "NULL frame");
    f->line(1)->v.store(seq, std::memory_order_release);
    return NT_OK;
}

extern "C" int nt_host_frame_wait_ack(nt_host_frame *f, uint32_t seq, uint32_t timeout_ms) {
    if (!f) return nt_fail_public(NT_ERR_INVALID, "NULL frame");
    if (wait_line(f->line(1), seq, timeout_ms) != NT_OK) return nt_fail_public(NT_ERR_TIMEOUT, "frame %u was not acknowledged within %u ms", seq, timeout_ms);
    return NT_OK;
}
