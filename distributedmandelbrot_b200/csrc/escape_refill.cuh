// escape_refill.cuh -- persistent-warp form of the escape-time kernel with LANE REFILL
// (mb_tile.flags & MB_TILE_LANE_REFILL; same counts as escape_tiles_kernel, bit for bit).
//
// Same numerics contract and the same deferred-test machinery as escape_kernels.cuh
// (DistributedMandelbrotWorkerCUDA.py:39-68: one pixel per thread at a time, blocks of K untested
// updates, exact replay of the block that tripped).  What changes is who owns a pixel for how long:
//
//   escape_tiles_kernel   a warp owns one 8x4 patch for its whole life; lanes that finish wait for
//                         the slowest lane of the patch -- at the set boundary one interior pixel
//                         (mrd updates) keeps 31 finished lanes idle (ncu: 28.4 of 32 lanes active
//                         on the cfg4 boundary block, profiles/r01_ncu_full_cfg4_boundary_4x4.txt).
//   escape_refill_kernel  warps are persistent.  All lanes run the same trip (two blocks of K
//                         untested updates) in lock step; after a trip one warp vote asks "did any lane
//                         trip, run out of updates or (shortcut) close a cycle?".  If not -- the only
//                         thing the interior ever executes -- the next trip starts: one VOTE and one
//                         branch per 2K updates.  If so, the lanes concerned replay / finish in a short
//                         cold section, store their result, and are REFILLED with the next pixels of the
//                         warp's patch stream; the other lanes keep their orbits.  Patches are handed
//                         out by one global atomic counter per launch (one atomicAdd per 32 pixels), so
//                         there is no per-CTA tail either: a warp only runs dry when the launch does.
//
// There is no tested "head": a fresh pixel starts with untested blocks at once (a pixel that
// escapes after 3 updates pays one trip plus one replay), so this form is meant for deep zooms where
// escape counts are in the hundreds or more (cfg4, cfg5); shallow frames keep escape_tiles_kernel.
// Lanes whose c lies in the |c| ~ 2 annulus (`risky`, see escape_kernels.cuh) or outside the plane
// are evaluated on the spot with a test after every update and never enter a deferred block.
#pragma once

#include "escape_kernels.cuh"

namespace mbk {

template <typename Real, typename OutT>
__device__ __forceinline__ void store_result(OutT *dst, int result, int lim)
{
    if constexpr (sizeof(OutT) == 4) *dst = result;
    else *dst = (OutT)count_to_u8((uint32_t)result, (uint32_t)lim + 1u);      // fused .py:96-98, mrd = lim + 1
}

// literal / fully tested evaluation of one pixel (risky annulus, exotic coordinates, mrd tails)
template <typename Real>
__device__ __forceinline__ int escape_tested(Real cr, Real ci, int lim)
{
    using O = Ops<Real>;
    Real x = cr, y = ci, xx = O::mul(x, x), yy = O::mul(y, y);
    if (!O::ordinary(cr, ci)) {
        for (int n = 1; n <= lim; ++n) {                 // .py:44-68 literally, IEEE compare, literal (2x)*y
            const Real a = O::sub(xx, yy);
            const Real b2 = O::mul(O::add(x, x), y);
            x = O::add(a, cr);
            y = O::add(b2, ci);
            xx = O::mul(x, x);
            yy = O::mul(y, y);
            if (O::ge4(O::add(xx, yy))) return n;
        }
        return 0;
    }
    for (int n = 1; n <= lim; ++n) {
        update<Real>(x, y, xx, yy, cr, ci);
        if (O::escaped(O::add(xx, yy))) return n;
    }
    return 0;
}

template <typename Real, typename OutT, int MIN_CTAS, bool PERIOD>
__global__ void __launch_bounds__(128, MIN_CTAS)
escape_refill_kernel(const __grid_constant__ TileArgs args, OutT *__restrict__ out)
{
    using O = Ops<Real>;
    constexpr int K = Defer<Real>::K;
    // shared-memory staging of the complex-plane bounds: one descriptor slot per warp, holding the
    // tile of the patch the warp is currently handing out
    __shared__ mb_tile s_desc[4];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    mb_tile &desc = s_desc[wid];
    const unsigned long long ppt = (unsigned long long)args.pcols * args.prows;   // patch slots per tile
    const unsigned long long total = ppt * args.n_tiles;

    // warp-uniform cursor over the warp's patch stream
    uint32_t pos = 32, pcol = 0, prow = 0;
    bool exhausted = false, canon = false;

    // per-lane pixel state (the orbit itself lives in the trip registers below)
    Real cr = 0, ci = 0, sx = 0, sy = 0;
    int n = 0, lim = 0;
    OutT *dst = nullptr;
    bool alive = false;
    bool pend = false;                 // PERIOD: orbit closed a cycle (result 0), store + refill still owed
    unsigned long long exec_acc = 0;
    // PERIOD: a lane that closes a cycle has nothing to replay, so its store and refill can wait until
    // kBatch lanes are free (or a lane needs the cold section anyway): one cold section serves them all
    constexpr unsigned kBatch = 4;

    // Give every lane with `want` the next pixel of the stream; (nx, ny) receive z0 = c.  Pixels that
    // must not enter deferred blocks are evaluated and stored right here and the lane asks again.
    auto refill = [&](bool want, Real &nx, Real &ny) {
        unsigned need = __ballot_sync(kFull, want);
        const bool whole = (need == kFull);
        bool all_plain = true;
        while (need) {
            if (pos >= 32u) {
                if (exhausted) break;
                unsigned long long g = 0;
                if (lane == 0) g = atomicAdd(args.counter, 1ull);
                g = __shfl_sync(kFull, g, 0);
                if (g >= total) { exhausted = true; break; }
                const uint32_t tile = (uint32_t)(g / ppt), p = (uint32_t)(g % ppt);
                prow = p / args.pcols;
                pcol = p - prow * args.pcols;
                __syncwarp();
                if (lane < sizeof(mb_tile) / 8) {
                    const mb_tile *src = args.n_inline ? &args.inl[tile] : args.dev + tile;
                    reinterpret_cast<uint64_t *>(&desc)[lane] = reinterpret_cast<const uint64_t *>(src)[lane];
                }
                __syncwarp();
                if (pcol * kWarpW >= desc.w || prow * kWarpH >= desc.h) continue;    // slot beyond this tile's own patch grid
                pos = 0;
            }
            const uint32_t rank = __popc(need & lt), avail = 32u - pos, cnt = __popc(need);
            const bool take = want && rank < avail;
            const uint32_t q = pos + rank;
            const bool fresh = whole && pos == 0 && cnt == 32u;      // a whole patch, lane i <- pixel i
            pos += cnt < avail ? cnt : avail;
            bool plain = false;
            if (take) {
                const uint32_t px = pcol * kWarpW + (q & 7u), py = prow * kWarpH + (q >> 3);
                if (px < desc.w && py < desc.h) {
                    const uint32_t d1 = desc.definition - 1u;
                    cr = O::coord(linspace_at(desc.x0 + px, d1, desc.start_r, desc.step_r, desc.stop_r));
                    ci = O::coord(linspace_at(desc.y0 + py, d1, desc.start_i, desc.step_i, desc.stop_i));
                    lim = (int)desc.mrd - 1;                         // updates n = 1 .. mrd-1 (.py:47)
                    dst = out + desc.out_offset + (size_t)py * desc.out_pitch + px;
                    const Real q4 = O::sub(O::add(O::mul(cr, cr), O::mul(ci, ci)), (Real)4);
                    const bool risky = (q4 <= Defer<Real>::kBand) && (q4 >= -Defer<Real>::kBand);
                    if (lim < 2 * K || risky || !O::ordinary(cr, ci)) {
                        const int r = lim > 0 ? escape_tested<Real>(cr, ci, lim) : 0;
                        store_result<Real, OutT>(dst, r, lim);
                        if constexpr (PERIOD) exec_acc += (unsigned long long)(r ? r : (lim > 0 ? lim : 0));
                    } else {
                        nx = cr; ny = ci;                            // z0 = c (.py:44-45)
                        sx = cr; sy = ci;
                        n = 0;
                        alive = true;
                        want = false;
                        plain = true;
                    }
                }
            }
            all_plain = all_plain && fresh && __all_sync(kFull, plain);
            need = __ballot_sync(kFull, want);
        }
        canon = whole && all_plain && (__ballot_sync(kFull, alive) == kFull);   // lane i holds pixel i of one patch: results can leave as 128-bit rows
    };

    Real x = 0, y = 0;
    refill(true, x, y);
    Real xx = O::mul(x, x), yy = O::mul(y, y);

    // Resolve the lanes that raised an event after a trip s -> m -> e, store and refill; on return
    // (ex, ey, exx, eyy) hold every lane's state for the next trip and n is up to date.
    auto cold = [&](Real sxx_, Real sy_, Real mx_, Real my_, Real mA, Real mB, Real &ex, Real &ey, Real &exx, Real &eyy) {
        bool fin = pend;                                             // parked since an earlier trip: result 0
        pend = false;
        int result = 0;
        if (alive) {
            const bool tA = Defer<Real>::tripped(mA), tB = Defer<Real>::tripped(mB);
            if (tA || tB) {
                // exact replay of the block that tripped: test after every update (.py:62-66)
                Real rx = tA ? sxx_ : mx_, ry = tA ? sy_ : my_;
                const int base = tA ? n : n + K;
                Real rxx = O::mul(rx, rx), ryy = O::mul(ry, ry);
                for (int j = 1; j <= K; ++j) {
                    update<Real>(rx, ry, rxx, ryy, cr, ci);
                    if (O::escaped(O::add(rxx, ryy))) { result = base + j; break; }
                }
                fin = true;
            } else {
                n += 2 * K;
                bool periodic = false;
                if constexpr (PERIOD) {
                    periodic = O::same_bits(ex, sx) && O::same_bits(ey, sy);
                    if (save_point(n)) { sx = ex; sy = ey; }
                }
                const int left = lim - n;
                if (periodic) {
                    fin = true;                                      // exact cycle: never escapes, result 0
                } else if (left < 2 * K) {
                    // fewer updates left than a trip: tested tail
                    Real rx = ex, ry = ey, rxx = exx, ryy = eyy;
                    for (int j = 1; j <= left; ++j) {
                        update<Real>(rx, ry, rxx, ryy, cr, ci);
                        if (O::escaped(O::add(rxx, ryy))) { result = n + j; break; }
                    }
                    if (result == 0) n = lim;
                    fin = true;
                }
            }
        }
        const unsigned finmask = __ballot_sync(kFull, fin);
        if (finmask == kFull && canon) {
            // the whole patch finished in the same trip (the interior): 128-bit row stores
            if constexpr (sizeof(OutT) == 4) {
                const int v1 = __shfl_down_sync(kFull, result, 1), v2 = __shfl_down_sync(kFull, result, 2);
                const int v3 = __shfl_down_sync(kFull, result, 3);
                const bool vec = (reinterpret_cast<uintptr_t>(dst - (lane & 3u)) & 15u) == 0;    // same answer in the 4 lanes of a group
                if (vec) { if ((lane & 3u) == 0) *reinterpret_cast<int4 *>(dst) = make_int4(result, v1, v2, v3); }
                else *dst = result;
            } else {
                const uint32_t b = count_to_u8((uint32_t)result, (uint32_t)lim + 1u);
                uint32_t v = b | (__shfl_down_sync(kFull, b, 1) << 8);
                v |= __shfl_down_sync(kFull, v, 2) << 16;
                const uint32_t hi = __shfl_down_sync(kFull, v, 4);
                const bool vec = (reinterpret_cast<uintptr_t>(dst - (lane & 7u)) & 7u) == 0;
                if (vec) { if ((lane & 7u) == 0) *reinterpret_cast<uint2 *>(dst) = make_uint2(v, hi); }
                else *dst = (OutT)b;
            }
        } else if (fin) {
            store_result<Real, OutT>(dst, result, lim);              // L2 merges the sector's other pixels
        }
        if (fin) {
            if constexpr (PERIOD) { if (alive) exec_acc += (unsigned long long)(result ? result : n); }   // parked lanes were counted when they parked
            alive = false;
        }
        Real nx = ex, ny = ey;
        refill(fin, nx, ny);
        if (fin) {                                                   // refilled (or dead: harmless values)
            ex = nx; ey = ny;
            exx = O::mul(nx, nx); eyy = O::mul(ny, ny);
        }
    };

    // after a trip: does any lane need the cold section now?  (PERIOD: lanes on a cycle are parked first)
    auto after_trip = [&](Real mA, Real mB, Real ex, Real ey) -> bool {
        bool hard = alive && (Defer<Real>::tripped(mA) || Defer<Real>::tripped(mB) || (lim - n < 4 * K));
        if constexpr (PERIOD) {
            if (alive && !hard && O::same_bits(ex, sx) && O::same_bits(ey, sy)) {
                alive = false;                                       // exact cycle: never escapes
                pend = true;
                exec_acc += (unsigned long long)(n + 2 * K);
            }
            const unsigned parked = __ballot_sync(kFull, pend);
            const unsigned live = __ballot_sync(kFull, alive);
            return __any_sync(kFull, hard) || __popc(parked) >= (int)kBatch || (parked && !live);
        } else {
            return __any_sync(kFull, hard);
        }
    };

    while (__any_sync(kFull, alive || pend)) {
        // ---- trip 1: (x, y) -> a -> b
        Real ax = x, ay = y, axx = xx, ayy = yy;
#pragma unroll
        for (int j = 0; j < K; ++j) update<Real>(ax, ay, axx, ayy, cr, ci);
        Real mA = O::add(axx, ayy);
        Real bx = ax, by = ay, bxx = axx, byy = ayy;
#pragma unroll
        for (int j = 0; j < K; ++j) update<Real>(bx, by, bxx, byy, cr, ci);
        Real mB = O::add(bxx, byy);
        if (after_trip(mA, mB, bx, by)) {
            cold(x, y, ax, ay, mA, mB, bx, by, bxx, byy);
            x = bx; y = by; xx = bxx; yy = byy;
            continue;
        }
        n += 2 * K;
        if constexpr (PERIOD) { if (save_point(n)) { sx = bx; sy = by; } }
        // ---- trip 2: b -> a -> (x, y)
        ax = bx; ay = by; axx = bxx; ayy = byy;
#pragma unroll
        for (int j = 0; j < K; ++j) update<Real>(ax, ay, axx, ayy, cr, ci);
        mA = O::add(axx, ayy);
        x = ax; y = ay; xx = axx; yy = ayy;
#pragma unroll
        for (int j = 0; j < K; ++j) update<Real>(x, y, xx, yy, cr, ci);
        mB = O::add(xx, yy);
        if (after_trip(mA, mB, x, y)) {
            cold(bx, by, ax, ay, mA, mB, x, y, xx, yy);
            continue;
        }
        n += 2 * K;
        if constexpr (PERIOD) { if (save_point(n)) { sx = x; sy = y; } }
    }

    if constexpr (PERIOD) {
        if (args.executed) {
            for (int o = 16; o > 0; o >>= 1) exec_acc += __shfl_down_sync(kFull, exec_acc, o);
            if (lane == 0 && exec_acc) atomicAdd(args.executed, exec_acc);
        }
    }
}

}  // namespace mbk
