// tile_wsq.cpp -- in-box tile dispatch: lock-free work-stealing index ranges in POSIX shm.
//
// Replaces, for the GPUs of ONE box, what the reference does over TCP with one chunk lease per
// connection: Distributer.TryGetNextNeededWorkload / HandleWorkloadRequest
// (DistributedMandelbrot/Distributer.cs:335-392).  The reference's scheme is pull-based
// (workers ask for the next chunk) and so is this one, but the "server" is a 64-byte slot per
// rank in a shared-memory segment and a claim is one compare-and-swap instead of a TCP
// round trip through a single-threaded accept loop (Distributer.cs:226-297).
//
// Layout: slot[r] = (head << 32) | tail, the half-open range of tile indices rank r still owns.
//   owner pop : CAS (h, t) -> (h + k, t), takes [h, h + k)
//   steal     : CAS victim (h, t) -> (h, t - k) with k = ceil((t - h) / 2), thief deposits
//               [t - k, t) into its own (empty) slot and pops from it.
// `remaining` counts indices not yet handed out for execution, so a rank that finds nothing to
// steal while a thief is mid-transfer waits instead of quitting early.
// Every index in [0, total) is handed out exactly once (ranges are disjoint and only shrink).
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mandelbrot_b200.h"

extern "C" int mb_internal_fail(int code, const char *msg);   // mandelbrot_capi.cu

namespace {

constexpr uint32_t kMagic = 0x4d425153u;   // "MBQS"
constexpr uint32_t kMaxRanks = 64;

struct alignas(64) Slot {
    uint64_t range;                         // (head << 32) | tail
    char pad[56];
};

struct Shared {
    uint32_t magic;
    uint32_t n_ranks;
    uint32_t total;
    uint32_t epoch;
    alignas(64) uint64_t remaining;
    alignas(64) Slot slot[kMaxRanks];
};

int qfail(int code, const char *fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return mb_internal_fail(code, buf);   // message readable through mb_last_error()
}

inline uint64_t pack(uint32_t h, uint32_t t) { return ((uint64_t)h << 32) | t; }
inline uint32_t head_of(uint64_t v) { return (uint32_t)(v >> 32); }
inline uint32_t tail_of(uint64_t v) { return (uint32_t)v; }

void arm(Shared *s, uint32_t total)
{
    const uint32_t n = s->n_ranks;
    for (uint32_t r = 0; r < n; ++r) {
        const uint32_t lo = (uint32_t)((uint64_t)total * r / n);
        const uint32_t hi = (uint32_t)((uint64_t)total * (r + 1) / n);
        __atomic_store_n(&s->slot[r].range, pack(lo, hi), __ATOMIC_RELAXED);
    }
    s->total = total;
    __atomic_store_n(&s->remaining, (uint64_t)total, __ATOMIC_RELAXED);
    __atomic_add_fetch(&s->epoch, 1u, __ATOMIC_RELEASE);
}

}  // namespace

struct mb_wsq {
    Shared *sh;
    int fd;
};

extern "C" {

int mb_wsq_create(const char *name, uint32_t n_ranks, uint32_t total, mb_wsq **out)
{
    if (!name || name[0] != '/' || !out) return qfail(MB_ERR_INVALID, "queue name must start with '/'");
    if (n_ranks == 0 || n_ranks > kMaxRanks) return qfail(MB_ERR_INVALID, "n_ranks must be in [1, %u]", kMaxRanks);
    shm_unlink(name);   // a stale segment from a crashed run must not leak old state
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return qfail(MB_ERR_SYS, "shm_open(%s): %s", name, strerror(errno));
    if (ftruncate(fd, sizeof(Shared)) != 0) {
        int e = errno;
        close(fd);
        shm_unlink(name);
        return qfail(MB_ERR_SYS, "ftruncate: %s", strerror(e));
    }
    void *p = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
        int e = errno;
        close(fd);
        shm_unlink(name);
        return qfail(MB_ERR_SYS, "mmap: %s", strerror(e));
    }
    Shared *s = static_cast<Shared *>(p);
    std::memset(s, 0, sizeof(Shared));
    s->n_ranks = n_ranks;
    arm(s, total);
    __atomic_store_n(&s->magic, kMagic, __ATOMIC_RELEASE);
    mb_wsq *q = static_cast<mb_wsq *>(std::malloc(sizeof(mb_wsq)));
    if (!q) return qfail(MB_ERR_NOMEM, "out of memory");
    q->sh = s;
    q->fd = fd;
    *out = q;
    return MB_OK;
}

int mb_wsq_open(const char *name, mb_wsq **out)
{
    if (!name || !out) return qfail(MB_ERR_INVALID, "NULL argument");
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return qfail(MB_ERR_SYS, "shm_open(%s): %s", name, strerror(errno));
    void *p = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
        int e = errno;
        close(fd);
        return qfail(MB_ERR_SYS, "mmap: %s", strerror(e));
    }
    Shared *s = static_cast<Shared *>(p);
    if (__atomic_load_n(&s->magic, __ATOMIC_ACQUIRE) != kMagic) {
        munmap(p, sizeof(Shared));
        close(fd);
        return qfail(MB_ERR_STATE, "segment %s is not an initialised tile queue", name);
    }
    mb_wsq *q = static_cast<mb_wsq *>(std::malloc(sizeof(mb_wsq)));
    if (!q) return qfail(MB_ERR_NOMEM, "out of memory");
    q->sh = s;
    q->fd = fd;
    *out = q;
    return MB_OK;
}

int mb_wsq_reset(mb_wsq *q, uint32_t total)
{
    if (!q) return qfail(MB_ERR_INVALID, "NULL queue");
    arm(q->sh, total);
    return MB_OK;
}

int mb_wsq_pop(mb_wsq *q, uint32_t rank, uint32_t want, uint32_t *first, uint32_t *got, uint32_t *stolen)
{
    if (!q || !first || !got) return qfail(MB_ERR_INVALID, "NULL argument");
    Shared *s = q->sh;
    if (rank >= s->n_ranks) return qfail(MB_ERR_INVALID, "rank %u >= n_ranks %u", rank, s->n_ranks);
    if (want == 0) want = 1;
    uint32_t was_stolen = 0;
    for (unsigned spins = 0;; ++spins) {
        // 1. pop from the head of my own range
        uint64_t cur = __atomic_load_n(&s->slot[rank].range, __ATOMIC_ACQUIRE);
        while (head_of(cur) < tail_of(cur)) {
            const uint32_t h = head_of(cur), t = tail_of(cur);
            const uint32_t k = (t - h < want) ? t - h : want;
            if (__atomic_compare_exchange_n(&s->slot[rank].range, &cur, pack(h + k, t), false,
                                            __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
                __atomic_sub_fetch(&s->remaining, (uint64_t)k, __ATOMIC_ACQ_REL);
                *first = h;
                *got = k;
                if (stolen) *stolen = was_stolen;
                return 1;
            }
        }
        // 2. my range is empty: steal the upper half of the fullest victim
        uint32_t best = rank, best_left = 0;
        for (uint32_t v = 0; v < s->n_ranks; ++v) {
            if (v == rank) continue;
            const uint64_t rv = __atomic_load_n(&s->slot[v].range, __ATOMIC_ACQUIRE);
            const uint32_t left = tail_of(rv) > head_of(rv) ? tail_of(rv) - head_of(rv) : 0;
            if (left > best_left) { best_left = left; best = v; }
        }
        if (best_left > 0) {
            uint64_t rv = __atomic_load_n(&s->slot[best].range, __ATOMIC_ACQUIRE);
            const uint32_t h = head_of(rv), t = tail_of(rv);
            if (h < t) {
                const uint32_t k = (t - h + 1) / 2;
                if (__atomic_compare_exchange_n(&s->slot[best].range, &rv, pack(h, t - k), false,
                                                __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
                    __atomic_store_n(&s->slot[rank].range, pack(t - k, t), __ATOMIC_RELEASE);
                    was_stolen = 1;
                }
            }
            continue;   // pop from my refilled range, or rescan after a lost race
        }
        // 3. nothing visible: drained, unless a thief is mid-transfer
        if (__atomic_load_n(&s->remaining, __ATOMIC_ACQUIRE) == 0) {
            *first = 0;
            *got = 0;
            if (stolen) *stolen = 0;
            return 0;
        }
        if (spins > 64) sched_yield();
    }
}

int64_t mb_wsq_remaining(mb_wsq *q)
{
    if (!q) return qfail(MB_ERR_INVALID, "NULL queue");
    return (int64_t)__atomic_load_n(&q->sh->remaining, __ATOMIC_ACQUIRE);
}

int mb_wsq_close(mb_wsq *q)
{
    if (!q) return MB_OK;
    munmap(q->sh, sizeof(Shared));
    close(q->fd);
    std::free(q);
    return MB_OK;
}

int mb_wsq_unlink(const char *name)
{
    if (!name) return qfail(MB_ERR_INVALID, "NULL name");
    if (shm_unlink(name) != 0 && errno != ENOENT) return qfail(MB_ERR_SYS, "shm_unlink(%s): %s", name, strerror(errno));
    return MB_OK;
}

}  // extern "C"
