/*
 * mandelbrot_b200.h -- C-ABI of the B200-native escape-time tile renderer.
 *
 * This is the drop-in boundary for ONE hot path of ofsouzap/DistributedMandelbrot: the
 * per-pixel z <- z^2 + c escape-time iteration over a DataChunk tile.  The reference exposes
 * no FFI of its own; the entry points below are what a binding for this path
 * would bind, each one citing the reference interface it replaces.  Paths are relative to the
 * reference root; ".py" = DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py.
 *
 * Conventions
 *   - extern "C", plain C types, no exceptions across the boundary, no torch types.
 *   - every function returns MB_OK (0) or a negative mb_status; mb_last_error() gives the
 *     message for the calling thread.
 *   - the caller owns every buffer; the library never frees caller memory.
 *   - `*_dev` pointers are device pointers valid on the device passed to mb_init() for the
 *     calling thread (e.g. torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as
 *     void* (NULL = the legacy default stream).  Device-pointer entry points only ENQUEUE
 *     work on `stream`; host-pointer entry points are synchronous.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails
 *     with MB_ERR_CUDA.
 */
#ifndef MANDELBROT_B200_H
#define MANDELBROT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 1

typedef enum mb_status {
    MB_OK = 0,
    MB_ERR_INVALID = -1,   /* bad argument (level == 0, index >= level, mrd out of range, ...) */
    MB_ERR_CUDA = -2,      /* CUDA runtime error / no device */
    MB_ERR_NOMEM = -3,
    MB_ERR_STATE = -4,     /* mb_init() not called on this thread's device */
    MB_ERR_SYS = -5        /* shm / OS error (tile queue) */
} mb_status;

/* arithmetic type of the iteration */
#define MB_DTYPE_F64 0     /* the reference's only type: int32(float64, float64, int32), .py:39 */
#define MB_DTYPE_F32 1     /* binary32 variant (no upstream counterpart) */

/* DataChunk constants (DataChunk.cs:14-15,20,27 ; .py:7-8,80) */
#define MB_MIN_AXIS (-2.0)
#define MB_MAX_AXIS (2.0)
#define MB_CHUNK_DEFINITION 4096u

/*
 * Tile descriptor: a sub-rectangle of one DataChunk's pixel grid.
 *
 * The six doubles are the chunk-level numpy.linspace parameters of .py:24-32 computed on the
 * host (mb_chunk_geometry) so that host and device share bits:
 *   v[k] = fl(fl(k*step) + start)  for k < definition-1 ;  v[definition-1] = stop.
 * Pixel (row, col) of the CHUNK has c = (r[col], i[row])  (.py:34-35: np.tile / np.repeat).
 * The tile covers chunk columns [x0, x0+w) and rows [y0, y0+h); its pixel (row, col) is
 * written to out[out_offset + (row-y0)*out_pitch + (col-x0)].  A tile never re-linspaces
 * itself: it indexes into the chunk-level grid.
 */
typedef struct mb_tile {
    double start_r, start_i;
    double stop_r, stop_i;
    double step_r, step_i;
    uint32_t definition;   /* samples per axis of the chunk (reference: 4096) */
    uint32_t mrd;          /* maximum recursion depth: updates n = 1 .. mrd-1 (.py:47) */
    uint32_t x0, y0, w, h;
    uint64_t out_offset;   /* in elements */
    uint32_t out_pitch;    /* elements between successive tile rows in out */
    uint32_t flags;        /* 0 or MB_TILE_* bits below; the same for every tile of one launch */
} mb_tile;                 /* 88 bytes */

/*
 * mb_tile.flags bit 0 -- exact orbit-periodicity shortcut (opt-in, off by default).  A lane whose
 * (x, y) state at a block end is BIT-EQUAL to an earlier block-end state is on an exact cycle of
 * the floating-point map; it can never escape, so the reference loop (.py:47-68) would run to
 * mrd-1 and return 0 -- the lane reports 0 at once.  Counts are bit-identical with the flag on
 * or off; only the number of EXECUTED updates drops (pixels inside the set dominate the deep
 * zooms).  Throughput measured with the flag is reported separately from the FP64-pipe roofline
 * numbers, which always count executed work.
 */
#define MB_TILE_PERIOD_SHORTCUT 1u
/*
 * mb_tile.flags bit 1 -- persistent warps with lane refill (escape_refill.cuh): lanes that finish
 * (escaped, out of updates, or on a cycle) are handed the next pixel instead of waiting for the
 * slowest lane of their 8x4 patch.  Bit-identical counts.  Meant for deep zooms (escape counts in
 * the hundreds or more): there is no tested head, a pixel that escapes after 3 updates pays a whole
 * trip of 16.  May be combined with MB_TILE_PERIOD_SHORTCUT.
 */
#define MB_TILE_LANE_REFILL 2u

/* ---- lifecycle ----------------------------------------------------------------------- */

/* Bind the calling thread to CUDA device `device` and create the library context for it. */
int mb_init(int device);
void mb_shutdown(void);
const char *mb_last_error(void);
int mb_abi_version(void);
/* number of CUDA devices, or a negative mb_status */
int mb_device_count(void);
/* SM count of the bound device (148 on B200) */
int mb_sm_count(void);
/*
 * Per-thread switch for the one-call entry points below (mb_process_workload_*): render their
 * chunk with MB_TILE_PERIOD_SHORTCUT (see mb_tile.flags; results are bit-identical either way).
 * Off by default.  Returns the previous setting.  Tile-array entry points take the flag per tile.
 */
int mb_set_period_shortcut(int enabled);
/* same switch for the whole flag word (MB_TILE_PERIOD_SHORTCUT | MB_TILE_LANE_REFILL); returns the
 * previous flags, or MB_ERR_INVALID for unknown bits */
int mb_set_tile_flags(uint32_t flags);
/*
 * Diagnostic for shortcut launches of THIS thread: counter_dev (device uint64, caller-zeroed, NULL
 * to disable) accumulates the updates the kernel actually executed, to be set against
 * mb_useful_iterations (the updates the reference semantics require).
 */
int mb_set_executed_counter(uint64_t *counter_dev);

/* ---- geometry (host only, no GPU needed) --------------------------------------------- */

/*
 * Replaces .py:75-78 (chunk_range / start_r / start_i) == DataChunk.GetLevelChunkRange
 * (DataChunk.cs:32-33) + StartValueReal/Imag (DataChunk.cs:59-66), and the linspace
 * parameters of gen_arrays (.py:24-32).  Every operation individually rounded in binary64.
 * Fills a whole-chunk tile (x0 = y0 = 0, w = h = definition, out_offset = 0,
 * out_pitch = definition).  Errors like the DataChunk ctor (DataChunk.cs:99-106):
 * level == 0, index >= level -> MB_ERR_INVALID.  definition must be >= 1, mrd <= INT32_MAX
 * (the reference kernel takes int32 mrd, .py:39).
 */
int mb_chunk_geometry(uint32_t level, uint32_t index_real, uint32_t index_imag,
                      uint32_t definition, uint32_t mrd, mb_tile *out);

/* one linspace sample of a tile axis, exactly as the device computes it (for tests) */
double mb_tile_coord(const mb_tile *t, int axis /*0 = real, 1 = imag*/, uint32_t k);

/* ---- the hot path --------------------------------------------------------------------- */

/*
 * Replaces calc_mb_value (.py:39-68) together with gen_arrays (.py:19-37): int32 escape
 * counts for `n_tiles` tiles, written to out_dev.  One CUDA thread per pixel; coordinates are
 * generated in-kernel from the descriptors (nothing but 88 bytes per tile is read).
 * Count semantics (.py:44-68): z0 = c; n = 1..mrd-1: z <- z^2 + c; first n with |z|^2 >= 4,
 * else 0.  The caller guarantees that every tile's output range
 * [out_offset, out_offset + (h-1)*out_pitch + w) lies inside out_dev and that tiles of one launch
 * do not overlap; tiles of one launch may differ in size, mrd and chunk.
 */
int mb_render_counts(const mb_tile *tiles, uint32_t n_tiles, int32_t *out_dev, int dtype,
                     void *stream);

/*
 * calc_mb_value exactly as the reference calls it at .py:92: explicit float64 coordinate
 * arrays (device pointers, n elements each) -> int32 counts.  One thread per element.
 */
int mb_calc_mb_value_dev(const double *r_dev, const double *i_dev, size_t n, uint32_t mrd,
                         int dtype, int32_t *out_dev, void *stream);
/* host-pointer convenience form of the same (copies in, runs, copies out; synchronous) */
int mb_calc_mb_value(const double *r_host, const double *i_host, size_t n, uint32_t mrd,
                     int dtype, int32_t *out_host);

/*
 * Same, with the uint8 scaling of .py:96-98 fused into the kernel epilogue:
 * u = uint8(ceil(float64(n) * 256 / mrd)), including the 256 -> 0 wrap.
 * This is the DataChunk payload (DataChunk.cs:20,27,77): one byte per pixel.
 */
int mb_render_u8(const mb_tile *tiles, uint32_t n_tiles, uint8_t *out_dev, int dtype,
                 void *stream);

/* Stand-alone uint8 epilogue (.py:96-98) over n device int32 counts. */
int mb_counts_to_u8(const int32_t *counts_dev, uint8_t *out_dev, size_t n, uint32_t mrd,
                    void *stream);

/*
 * Metric helper: sum over pixels of (n ? n : mrd-1) = pixel-iterations the reference
 * semantics require.  Accumulates INTO *sum_dev (a device uint64 the caller
 * zeroes); asynchronous on `stream`.
 */
int mb_useful_iterations(const int32_t *counts_dev, size_t n, uint32_t mrd, uint64_t *sum_dev,
                         void *stream);

/*
 * Chunk classification used by the server on arrival (DataChunk.IsNeverChunk /
 * IsImmediateChunk, DataChunk.cs:82,87): flags_dev[0] = 1 iff every byte == 0,
 * flags_dev[1] = 1 iff every byte == 1.  Asynchronous on `stream`.
 */
int mb_classify_u8(const uint8_t *payload_dev, size_t n, uint32_t *flags_dev, void *stream);

/* ---- DataChunk serialisation on the device ---------------------------------- */
/*
 * Replaces DataChunk.Serialize (DataChunk.cs:173-206) with its two serializers
 * (DataChunkSerializer.cs:29-49 raw, :51-100 RLE) for a BATCH of equally sized chunks in one set
 * of launches: chunk c is read from payloads_dev + c*payload_stride (n bytes) and written to
 * out_dev + c*out_stride as 1 codec byte (0x00 raw / 0x01 RLE) + either the n payload bytes or
 * one (uint32 LE runLength, uint8 value) record per maximal run; RLE only when strictly shorter.
 * info_dev[c] receives the serialised length, the run count and the IndexEntry type the server
 * would derive (DataChunk.cs:82,87; DataStorage.cs:71-84).  out_stride >= n + 1, a multiple of 16
 * for speed (mb_serialize_out_stride).  n < 2^32, n_chunks <= 65535.  Asynchronous on `stream`.
 */
typedef struct mb_chunk_info {
    uint64_t out_len;      /* serialised bytes, codec byte included */
    uint32_t runs;
    uint32_t codec;        /* 0 raw, 1 RLE */
    uint32_t entry_type;   /* 0 Regular, 1 Never (all 0), 2 Immediate (all 1) */
    uint32_t first_value;
    uint32_t reserved[2];
} mb_chunk_info;           /* 32 bytes */

size_t mb_serialize_out_stride(size_t n);
size_t mb_serialize_workspace_bytes(size_t n, uint32_t n_chunks);
int mb_serialize_chunks(const uint8_t *payloads_dev, size_t n, uint32_t n_chunks,
                        size_t payload_stride, uint8_t *out_dev, size_t out_stride,
                        void *workspace_dev, size_t workspace_bytes, mb_chunk_info *info_dev,
                        void *stream);
/* host-pointer form for one chunk (copies in, serialises on the device, copies out; synchronous) */
int mb_serialize_chunk_host(const uint8_t *payload_host, size_t n, uint8_t *out_host,
                            size_t out_capacity, uint64_t *out_len);

/*
 * process_workload + what the server does with the result on arrival, in one call and without
 * the 16 MiB round trip: render (.py:70-100), classify (DataChunk.cs:82,87 ->
 * IndexEntry type, DataStorage.cs:71-84: 0 Regular, 1 Never, 2 Immediate) and serialise
 * (DataChunk.cs:173-206) on the device; only the serialised bytes come back.
 * out_host: definition^2 + 1 bytes.  For Never/Immediate chunks the reference stores no file
 * (DataStorage.cs:424-425): *out_len is still set, callers may ignore the bytes.
 */
int mb_process_workload_serialized(uint32_t level, uint32_t mrd, uint32_t index_real,
                                   uint32_t index_imag, uint32_t definition, uint8_t *out_host,
                                   size_t out_capacity, uint64_t *out_len, int32_t *entry_type);

/* Page-locked host memory for payload buffers (cudaHostAlloc / cudaFreeHost): a device->host copy
 * into it runs at PCIe speed without the driver's staging pass that pageable memory needs. */
int mb_host_alloc(size_t bytes, void **host_ptr);
int mb_host_free(void *host_ptr);

/* ---- one-call equivalents of process_workload (host buffers, synchronous) ------------- */

/*
 * Replaces process_workload(level, mrd, index_real, index_imag) -> uint8[definition^2]
 * (.py:70-100); the reference fixes definition = 4096 (.py:80).  host_out may be pageable or
 * pinned memory of definition^2 bytes.  The device->host copy is inside the call.
 */
int mb_process_workload_u8(uint32_t level, uint32_t mrd, uint32_t index_real,
                           uint32_t index_imag, uint32_t definition, uint8_t *host_out);

/*
 * The worker loop's compute for a LIST of workloads in one call (.py:178-184 calling .py:70-100):
 * workloads = n_workloads records of 4 uint32 (level, mrd, indexReal, indexImag -- the wire order,
 * DistributerWorkload.cs:53-77); chunk k's payload lands at host_out + k * definition^2.  Small
 * chunks are grouped into launches of about launch_bytes (0 = 4 MiB) of payload, launches and their
 * device->host copies are pipelined internally; synchronous as a whole.  host_out: page-locked
 * memory (mb_host_alloc) lets the copies overlap the kernels.  Honours mb_set_tile_flags.
 */
int mb_process_workloads_u8(const uint32_t *workloads, uint32_t n_workloads, uint32_t definition,
                            uint8_t *host_out, size_t launch_bytes);

/* same, int32 counts (what calc_mb_value returns before .py:96), definition^2 * 4 bytes */
int mb_process_workload_counts(uint32_t level, uint32_t mrd, uint32_t index_real,
                               uint32_t index_imag, uint32_t definition, int dtype,
                               int32_t *host_out);

/* ---- in-box tile dispatch: lock-free work-stealing ranges in POSIX shared memory ------- */
/*
 * Replaces Distributer.TryGetNextNeededWorkload / HandleWorkloadRequest
 * (Distributer.cs:335-392) for the in-box path: instead of leasing chunks over TCP, the
 * `n_ranks` GPU processes of one box share a shm segment holding one [head, tail) range of
 * tile indices per rank (a static contiguous split of [0, total)).  A rank pops from the head
 * of its own range; when that is empty it steals the upper half of the fullest victim's range.
 * All operations are single 64-bit CAS; every index in [0, total) is handed out exactly once.
 * Host-only: usable without a GPU.
 */
typedef struct mb_wsq mb_wsq;

/* rank 0: create + initialise (name like "/mbq_1234"); others: open after a barrier */
int mb_wsq_create(const char *name, uint32_t n_ranks, uint32_t total, mb_wsq **out);
int mb_wsq_open(const char *name, mb_wsq **out);
/* re-arm the same segment for another pass over [0, total) (call on one rank, between barriers) */
int mb_wsq_reset(mb_wsq *q, uint32_t total);
/*
 * Claim up to `want` consecutive indices for `rank`.  Returns 1 and sets [*first, *first + *got)
 * when work was obtained (*stolen = 1 if it came from another rank's range), 0 when the whole
 * queue is drained, negative mb_status on error.
 */
int mb_wsq_pop(mb_wsq *q, uint32_t rank, uint32_t want, uint32_t *first, uint32_t *got,
               uint32_t *stolen);
/* indices not yet handed out (all ranks); lets callers shrink their claims near the end */
int64_t mb_wsq_remaining(mb_wsq *q);
int mb_wsq_close(mb_wsq *q);
int mb_wsq_unlink(const char *name);

/*
 * The per-GPU dispatch loop of the in-box frame renderer, in native code (a Python loop costs
 * ~60 us per claim, which shows once a band of exterior pixels renders in 10 us): claim up to
 * `bands_per_claim` consecutive tiles for `rank` (shrinking to single tiles near the end:
 * guided self-scheduling), launch them with mb_render_counts on `stream`, keep at most `depth`
 * launches in flight, repeat until the queue is drained.  claimed[2*k], claimed[2*k+1] receive
 * the (first, count) of claim k (capacity max_claims pairs); *n_claims, *n_stolen count them.
 * Returns after the last launch is ENQUEUED (not finished).
 */
int mb_frame_render_stolen(mb_wsq *q, uint32_t rank, uint32_t n_ranks, const mb_tile *tiles,
                           uint32_t n_tiles, uint32_t bands_per_claim, uint32_t depth,
                           int32_t *out_dev, int dtype, void *stream, uint32_t *claimed,
                           uint32_t max_claims, uint32_t *n_claims, uint32_t *n_stolen);

/* wall-clock (CLOCK_MONOTONIC, seconds) of the calling thread's last dispatch loop: when its first and
 * its last claim were taken -- with the caller's own end-of-step time this gives the queue-drain
 * tail (last claim -> last kernel end) that names the scaling limiter (bench.py) */
int mb_frame_last_stats(double *first_claim_s, double *last_claim_s);

/* ---- one frame of DataChunk payloads in shared, page-locked host memory ----------------------- */
/*
 * End-to-end form of the in-box path: the uint8 payloads of a frame (what process_workload
 * returns per chunk, .py:70-100; DataChunk.cs:20,27) live ONCE in a POSIX shm segment that every
 * GPU process of the box maps and page-locks.  The unit of dispatch is a row band ("unit"), so one
 * expensive chunk is rendered by several GPUs; every band is copied device->host straight to its
 * final offset of its chunk ("group" = `units_per_group` consecutive units) -- no gather, no
 * second copy.  Replaces the per-chunk TCP return of Distributer.cs:400-458 inside one box.
 * unit_done / group_left counters in the segment tell which payloads are complete and which
 * bands a dead rank left behind (mb_hostframe_missing; the reference's answer is its lease
 * timeout, Distributer.cs:22,317-330).  Host-only except mb_hostframe_pin/unpin.
 */
typedef struct mb_hostframe mb_hostframe;
int mb_hostframe_create(const char *name, uint64_t payload_bytes, uint32_t n_units,
                        uint32_t units_per_group, mb_hostframe **out);   /* one rank; name like "/mbf_1" */
int mb_hostframe_open(const char *name, mb_hostframe **out);              /* the others, after a barrier */
uint8_t *mb_hostframe_data(mb_hostframe *f);                              /* payload area of this mapping */
uint64_t mb_hostframe_bytes(mb_hostframe *f);
uint32_t mb_hostframe_units(mb_hostframe *f);
uint32_t mb_hostframe_groups(mb_hostframe *f);
int mb_hostframe_reset(mb_hostframe *f);                                  /* re-arm for another pass (one rank, between barriers) */
/* re-arm with only units [0, active_units) taking part (the rest counts as delivered): a frame sized for
 * B chunks may carry a shorter last batch */
int mb_hostframe_reset_active(mb_hostframe *f, uint32_t active_units);
/* page-lock / unlock THIS process's mapping on the bound device (cudaHostRegister) */
int mb_hostframe_pin(mb_hostframe *f);
int mb_hostframe_unpin(mb_hostframe *f);
/*
 * Publish units [first_unit, first_unit + count) as landed.  Groups whose last unit this was are
 * appended to completed[] (capacity max_completed); *n_completed is incremented per such group
 * (it is NOT reset: pass a zeroed counter).  Idempotent per unit.
 */
int mb_hostframe_mark_done(mb_hostframe *f, uint32_t first_unit, uint32_t count, uint32_t *completed,
                           uint32_t max_completed, uint32_t *n_completed);
int64_t mb_hostframe_group_left(mb_hostframe *f, uint32_t group);         /* units still missing */
/* units not (yet) landed: after the queue is drained and every live rank has returned, these were
 * claimed by a rank that died; units[] may be NULL to only count */
int mb_hostframe_missing(mb_hostframe *f, uint32_t *units, uint32_t max_units, uint32_t *n_missing);
int mb_hostframe_close(mb_hostframe *f);
int mb_hostframe_unlink(const char *name);

/*
 * Per-GPU dispatch loop of the end-to-end in-box path: like mb_frame_render_stolen, but every
 * claim is rendered with the fused uint8 epilogue (.py:96-98) into a device staging ring and copied
 * to tiles[k].out_offset (BYTES) of the host frame; tile k is unit k, its rows must be contiguous
 * (out_pitch == w).  Returns when every band this rank claimed has LANDED in host memory and is
 * published (mb_hostframe_mark_done); completed[] receives the chunks this rank completed.
 */
int mb_frame_render_u8_host(mb_wsq *q, uint32_t rank, uint32_t n_ranks, const mb_tile *tiles,
                            uint32_t n_tiles, uint32_t bands_per_claim, uint32_t depth,
                            mb_hostframe *frame, int dtype, uint32_t *claimed, uint32_t max_claims,
                            uint32_t *n_claims, uint32_t *n_stolen, uint32_t *completed,
                            uint32_t max_completed, uint32_t *n_completed);
/* same for an explicit list of units (repairing what mb_hostframe_missing reported) */
int mb_frame_render_u8_host_units(const uint32_t *units, uint32_t n_units, const mb_tile *tiles,
                                  uint32_t n_tiles, uint32_t depth, mb_hostframe *frame, int dtype,
                                  uint32_t *completed, uint32_t max_completed, uint32_t *n_completed);

/* ---- frame buffer shared between the GPU processes of one box (peer memory over NVLink) ------- */
/*
 * Alternative to the final gather: rank 0 allocates the frame buffer (mb_peer_alloc) and exports
 * a 64-byte handle; the other ranks open it (mb_peer_open) and pass the returned pointer as
 * out_dev to mb_render_counts / mb_frame_render_stolen, so the kernel's 128-bit count stores land
 * directly in rank 0's HBM over NVLink/NVSwitch -- the "collective" is fused into the kernel's
 * epilogue and nothing is left to gather.  (CUDA IPC; same box only.)
 */
#define MB_PEER_HANDLE_BYTES 64
int mb_peer_alloc(size_t bytes, void **dev_ptr, void *handle_out);
int mb_peer_open(const void *handle, void **dev_ptr);
int mb_peer_close(void *dev_ptr);     /* for pointers from mb_peer_open */
int mb_peer_free(void *dev_ptr);      /* for pointers from mb_peer_alloc */

/* ---- roofline denominators, measured on the bound device ------------------------------ */
/*
 * Dependency-free DFMA (dtype F64) / FFMA (F32) issue-rate microbenchmark: every SM runs
 * `chains` independent FMA chains per thread for `iters` steps.  *tera_fma_per_s = FMA-pipe
 * warp-lane operations per second / 1e12 (multiply by 2 for FLOP/s).  Synchronous.
 */
int mb_measure_fma_peak(int dtype, int iters, double *tera_fma_per_s, float *elapsed_ms);
/*
 * Same harness, selectable instruction mix: mode 0 = independent FMA chains (as above),
 * mode 1 = the 7-instruction Mandelbrot update (3 MUL, 3 ADD/SUB, 1 FMA) with no escape test
 * and no divergence -- the arithmetic ceiling of the escape loop itself.
 * *tera_ops_per_s counts FMA-pipe lane operations per second / 1e12.
 */
int mb_measure_pipe_rate(int dtype, int mode, int iters, double *tera_ops_per_s,
                         float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* MANDELBROT_B200_H */
