// host_frame.cpp -- one frame of DataChunk payloads in POSIX shared memory, filled band by band
// by the GPU processes of one box (in-box path).
//
// What it replaces: in the reference every worker holds its finished 16 MiB payload in private
// memory and ships it to the single-threaded Distributer over its own TCP connection
// (DistributedMandelbrotWorkerCUDA.py:148-172 ; Distributer.cs:400-458), which copies it once
// more into a DataChunk.  Inside one box the payloads of a frame live ONCE, in a shm segment every
// GPU process maps and page-locks (mb_hostframe_pin): each rank's device->host copies land at the
// band's final offset of its chunk, so there is no gather and no second copy, and a chunk may be
// rendered by several GPUs at once (the unit of dispatch is a row band, not a chunk).
//
// Completion accounting, also in the segment, all lock-free:
//   unit_done[u]   1 once band u has landed in host memory (set by the rank that rendered it)
//   group_left[g]  bands of chunk g still missing; the rank whose decrement reaches 0 "completes"
//                  the chunk (its payload is now exactly what process_workload returns) and reports
//                  it to its caller, e.g. to send it to a Distributer or write it to the store.
// Fault handling (the reference's answer is the lease timeout, Distributer.cs:22,317-330): a band
// that was claimed from the tile queue by a rank that died never gets its unit_done flag;
// mb_hostframe_missing lists such bands once the queue is drained so that a surviving rank can
// render them again (mb_frame_render_u8_host_units) instead of publishing an incomplete frame.
#include <fcntl.h>
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

constexpr uint32_t kMagic = 0x4d424846u;   // "MBHF"
constexpr size_t kAlign = 4096;

struct Header {
    uint32_t magic;
    uint32_t n_units;
    uint32_t n_groups;
    uint32_t units_per_group;
    uint64_t payload_bytes;
    uint64_t data_offset;      // page-aligned start of the payload area
    uint64_t total_bytes;
    uint32_t epoch;
    uint32_t pad;
    // uint32_t group_left[n_groups] ; uint8_t unit_done[n_units] follow
};

int hfail(int code, const char *fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return mb_internal_fail(code, buf);
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct mb_hostframe {
    Header *h;
    uint32_t *group_left;
    uint8_t *unit_done;
    uint8_t *data;
    size_t total_bytes;
    int fd;
    int pinned;
};

namespace {

void bind(mb_hostframe *f, void *base)
{
    f->h = static_cast<Header *>(base);
    f->group_left = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(base) + sizeof(Header));
    f->unit_done = reinterpret_cast<uint8_t *>(f->group_left + f->h->n_groups);
    f->data = static_cast<uint8_t *>(base) + f->h->data_offset;
    f->total_bytes = f->h->total_bytes;
}

// units [0, active) take part in the next pass; the rest counts as delivered (a frame sized for B chunks
// may carry fewer: the last batch of leases a Distributer hands out)
void arm(mb_hostframe *f, uint32_t active)
{
    const Header *h = f->h;
    if (active > h->n_units) active = h->n_units;
    for (uint32_t g = 0; g < h->n_groups; ++g) {
        const uint32_t lo = g * h->units_per_group;
        const uint32_t hi = lo + h->units_per_group < h->n_units ? lo + h->units_per_group : h->n_units;
        const uint32_t cnt = active > lo ? (active < hi ? active - lo : hi - lo) : 0u;
        __atomic_store_n(&f->group_left[g], cnt, __ATOMIC_RELAXED);
    }
    std::memset(f->unit_done, 0, active);
    std::memset(f->unit_done + active, 1, h->n_units - active);
    __atomic_add_fetch(&f->h->epoch, 1u, __ATOMIC_RELEASE);
}

}  // namespace

extern "C" {

int mb_hostframe_create(const char *name, uint64_t payload_bytes, uint32_t n_units, uint32_t units_per_group,
                        mb_hostframe **out)
{
    if (!name || name[0] != '/' || !out) return hfail(MB_ERR_INVALID, "frame name must start with '/'");
    if (payload_bytes == 0 || n_units == 0 || units_per_group == 0) return hfail(MB_ERR_INVALID, "empty frame");
    const uint32_t n_groups = (n_units + units_per_group - 1) / units_per_group;
    const size_t meta = sizeof(Header) + sizeof(uint32_t) * (size_t)n_groups + n_units;
    const size_t data_offset = round_up(meta, kAlign);
    const size_t total = data_offset + round_up((size_t)payload_bytes, kAlign);
    shm_unlink(name);
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return hfail(MB_ERR_SYS, "shm_open(%s): %s", name, strerror(errno));
    if (ftruncate(fd, (off_t)total) != 0) {
        int e = errno;
        close(fd);
        shm_unlink(name);
        return hfail(MB_ERR_SYS, "ftruncate(%zu bytes): %s", total, strerror(e));
    }
    void *p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
    if (p == MAP_FAILED) {
        int e = errno;
        close(fd);
        shm_unlink(name);
        return hfail(MB_ERR_SYS, "mmap(%zu bytes): %s", total, strerror(e));
    }
    mb_hostframe *f = static_cast<mb_hostframe *>(std::calloc(1, sizeof(mb_hostframe)));
    if (!f) { munmap(p, total); close(fd); shm_unlink(name); return hfail(MB_ERR_NOMEM, "out of memory"); }
    Header *h = static_cast<Header *>(p);
    h->n_units = n_units;
    h->n_groups = n_groups;
    h->units_per_group = units_per_group;
    h->payload_bytes = payload_bytes;
    h->data_offset = data_offset;
    h->total_bytes = total;
    h->epoch = 0;
    bind(f, p);
    f->fd = fd;
    arm(f, n_units);
    __atomic_store_n(&h->magic, kMagic, __ATOMIC_RELEASE);
    *out = f;
    return MB_OK;
}

int mb_hostframe_open(const char *name, mb_hostframe **out)
{
    if (!name || !out) return hfail(MB_ERR_INVALID, "NULL argument");
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return hfail(MB_ERR_SYS, "shm_open(%s): %s", name, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(Header)) {
        close(fd);
        return hfail(MB_ERR_STATE, "segment %s is too small to be a host frame", name);
    }
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
        int e = errno;
        close(fd);
        return hfail(MB_ERR_SYS, "mmap: %s", strerror(e));
    }
    const Header *h = static_cast<const Header *>(p);
    if (__atomic_load_n(&h->magic, __ATOMIC_ACQUIRE) != kMagic || h->total_bytes != (uint64_t)st.st_size) {
        munmap(p, (size_t)st.st_size);
        close(fd);
        return hfail(MB_ERR_STATE, "segment %s is not an initialised host frame", name);
    }
    mb_hostframe *f = static_cast<mb_hostframe *>(std::calloc(1, sizeof(mb_hostframe)));
    if (!f) { munmap(p, (size_t)st.st_size); close(fd); return hfail(MB_ERR_NOMEM, "out of memory"); }
    bind(f, p);
    f->fd = fd;
    *out = f;
    return MB_OK;
}

uint8_t *mb_hostframe_data(mb_hostframe *f) { return f ? f->data : nullptr; }
uint64_t mb_hostframe_bytes(mb_hostframe *f) { return f ? f->h->payload_bytes : 0; }
uint32_t mb_hostframe_units(mb_hostframe *f) { return f ? f->h->n_units : 0; }
uint32_t mb_hostframe_groups(mb_hostframe *f) { return f ? f->h->n_groups : 0; }

int mb_hostframe_reset(mb_hostframe *f)
{
    if (!f) return hfail(MB_ERR_INVALID, "NULL frame");
    arm(f, f->h->n_units);
    return MB_OK;
}

int mb_hostframe_reset_active(mb_hostframe *f, uint32_t active_units)
{
    if (!f) return hfail(MB_ERR_INVALID, "NULL frame");
    if (active_units > f->h->n_units) return hfail(MB_ERR_INVALID, "%u active units, the frame has %u", active_units, f->h->n_units);
    arm(f, active_units);
    return MB_OK;
}

int mb_hostframe_mark_done(mb_hostframe *f, uint32_t first_unit, uint32_t count, uint32_t *completed,
                           uint32_t max_completed, uint32_t *n_completed)
{
    if (!f) return hfail(MB_ERR_INVALID, "NULL frame");
    const Header *h = f->h;
    if ((uint64_t)first_unit + count > h->n_units) return hfail(MB_ERR_INVALID, "units [%u, %u) outside the frame", first_unit, first_unit + count);
    uint32_t nc = n_completed ? *n_completed : 0;
    for (uint32_t u = first_unit; u < first_unit + count; ++u) {
        // the band's bytes were written by a DMA that has completed; publish them before the flag
        if (__atomic_exchange_n(&f->unit_done[u], (uint8_t)1, __ATOMIC_ACQ_REL) != 0) continue;   // repaired twice: count once
        const uint32_t g = u / h->units_per_group;
        if (__atomic_sub_fetch(&f->group_left[g], 1u, __ATOMIC_ACQ_REL) == 0) {
            if (completed && nc < max_completed) completed[nc] = g;
            ++nc;
        }
    }
    if (n_completed) *n_completed = nc;
    return MB_OK;
}

int64_t mb_hostframe_group_left(mb_hostframe *f, uint32_t group)
{
    if (!f || group >= f->h->n_groups) return hfail(MB_ERR_INVALID, "bad group");
    return (int64_t)__atomic_load_n(&f->group_left[group], __ATOMIC_ACQUIRE);
}

int mb_hostframe_missing(mb_hostframe *f, uint32_t *units, uint32_t max_units, uint32_t *n_missing)
{
    if (!f || !n_missing) return hfail(MB_ERR_INVALID, "NULL argument");
    uint32_t n = 0;
    for (uint32_t u = 0; u < f->h->n_units; ++u) {
        if (__atomic_load_n(&f->unit_done[u], __ATOMIC_ACQUIRE)) continue;
        if (units && n < max_units) units[n] = u;
        ++n;
    }
    *n_missing = n;
    return MB_OK;
}

// used by mandelbrot_capi.cu (page-locking needs the CUDA runtime, which this file does not link)
int mb_internal_hostframe_region(mb_hostframe *f, void **base, size_t *bytes, int **pinned_flag)
{
    if (!f) return MB_ERR_INVALID;
    *base = f->data;
    *bytes = round_up((size_t)f->h->payload_bytes, kAlign);
    *pinned_flag = &f->pinned;
    return MB_OK;
}

int mb_hostframe_close(mb_hostframe *f)
{
    if (!f) return MB_OK;
    if (f->pinned) return hfail(MB_ERR_STATE, "mb_hostframe_unpin before mb_hostframe_close");
    munmap(f->h, f->total_bytes);
    close(f->fd);
    std::free(f);
    return MB_OK;
}

int mb_hostframe_unlink(const char *name)
{
    if (!name) return hfail(MB_ERR_INVALID, "NULL name");
    if (shm_unlink(name) != 0 && errno != ENOENT) return hfail(MB_ERR_SYS, "shm_unlink(%s): %s", name, strerror(errno));
    return MB_OK;
}

}  // extern "C"
