// Streaming multi-step kernel for the lossless Acoustic2D leapfrog (pyfds/acoustics.py:111-128):
// K time steps per launch with ONE read and ONE write of the state (temporal blocking).
//
// Work decomposition -- every warp is autonomous, there is no block-level synchronisation in the loop:
//   * a warp owns an x-strip of 120 cells and a chunk of rows; it streams the strip (128 cells wide:
//     4 halo cells either side) row by row through a K-stage time pipeline held in REGISTERS.
//     Stage t receives row q at time level t and emits row q-1 at level t+1, so after K stages the
//     warp stores row r-K at level K while row r is being read. Per stage only three row-fragments
//     stay live (p after boundaries, new vx, new vy of the previous row): 24 registers per lane.
//   * each lane holds 4 consecutive cells; the x-neighbours that live in the adjacent lane travel by
//     warp shuffle (p to the right-hand lane for the backward difference, new vx to the left-hand
//     lane for the forward difference). The outermost lanes compute garbage that never reaches the
//     120 owned cells: the dependency cone grows one cell per step and K <= 4.
//   * rows are fetched by the bulk-copy engine (TMA, cp.async.bulk + mbarrier) into a per-warp ring of
//     shared-memory slots, kRingDepth rows ahead of the arithmetic, so DRAM latency is hidden without
//     spending registers on loads in flight. Strips are addressed by flat cell index, which gives the
//     reference's row wrap (pyfds/fields.py:290-297) and the zero padding at the grid ends for free.
//   * material coefficients come from a 4 x 64 table in shared memory; rows whose 128 cells share one
//     material (almost all) take a warp-uniform fast path with the four coefficients in registers.
//   * boundary operations, sources and probes are applied inside the stages at the right time level
//     (per-cell flag bits in the material map byte, slow path only where a flag is set).
//
// Arithmetic is the same sequence of __dmul_rn/__dadd_rn as the one-step kernel, so results are
// bitwise identical to it and to the reference (tests/test_gpu_parity.py).
#pragma once

#include "fds_common.cuh"

namespace fds {

constexpr int kStripCells = 128;     // cells a warp streams per row (4 per lane)
constexpr int kStripHalo = 4;        // halo cells either side = one lane
constexpr int kStripStride = kStripCells - 2 * kStripHalo;   // 120 owned cells per strip
constexpr int kRingDepth = 6;        // rows in flight per warp
constexpr int kStreamWarps = 4;      // warps per CTA
constexpr int kMaxStreamSteps = 4;   // K
constexpr int kSlotBytes = 3 * kStripCells * 8 + 160;         // p, vx, vy rows + 144 map bytes (padded)
constexpr int kWarpRingBytes = kRingDepth * kSlotBytes + 64;  // + mbarriers

struct Stream2DArgs {
    const double *in[3];
    double *out[3];
    long long nx;          // row length, multiple of 4
    long long row_begin;   // rows [row_begin, row_end) are produced
    long long row_end;
    int chunk_rows;        // rows per task
    int n_strips;
    long long n_tasks;
    long long sig_index;   // first step - sig_first_step
    long long ring_row;    // probe record of the first step
};

__device__ __forceinline__ unsigned smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA without a tensor map)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, void *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

__device__ __forceinline__ double shfl_up1(double v) {
    return __shfl_up_sync(0xffffffffu, v, 1);
}
__device__ __forceinline__ double shfl_down1(double v) {
    return __shfl_down_sync(0xffffffffu, v, 1);
}

// Row metadata that travels through the pipeline with the row.
struct RowInfo {
    unsigned ids;      // 4 map bytes of this lane's cells
    int uniform;       // material id shared by all 128 cells of the row, or -1
    bool any_bound;    // some cell of the row (any lane) carries a boundary operation
    bool any_probe;
};

template <int K>
__global__ void __launch_bounds__(kStreamWarps * 32, 2) stream2d_kernel(Stream2DArgs a, StepTables t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double tabs[4][kMaxMaterials];   // FDS_TAB_GX, GY, FX, FY

    for (int k = threadIdx.x; k < 4 * kMaxMaterials; k += blockDim.x) (&tabs[0][0])[k] = t.tab[k];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * kStreamWarps + warp;
    if (task >= a.n_tasks) return;
    const long long nx = a.nx;
    const int strip = (int)(task % a.n_strips);
    const long long chunk = task / a.n_strips;
    const long long ys = a.row_begin + chunk * a.chunk_rows;
    const long long ye = min(ys + (long long)a.chunk_rows, a.row_end);
    const long long xs = (long long)strip * kStripStride - kStripHalo;   // column of lane 0, cell 0
    const long long r0 = ys - K, r1 = ye + K;                            // rows streamed in

    unsigned char *ring = smem_raw + warp * kWarpRingBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + kRingDepth * kSlotBytes);

    if (lane == 0) {
        for (int d = 0; d < kRingDepth; ++d) mbar_init(&bars[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](long long r, int slot) {
        const long long base = r * nx + xs;             // flat cell index of the strip start
        const long long base16 = base & ~15LL;          // map bytes: 16-byte aligned window
        unsigned char *dst = ring + slot * kSlotBytes;
        mbar_expect_tx(&bars[slot], 3 * kStripCells * 8 + 144);
        bulk_load(dst, a.in[0] + base, kStripCells * 8, &bars[slot]);
        bulk_load(dst + kStripCells * 8, a.in[1] + base, kStripCells * 8, &bars[slot]);
        bulk_load(dst + 2 * kStripCells * 8, a.in[2] + base, kStripCells * 8, &bars[slot]);
        bulk_load(dst + 3 * kStripCells * 8, t.map + base16, 144, &bars[slot]);
    };
    if (lane == 0)
        for (int d = 0; d < kRingDepth && r0 + d < r1; ++d) issue(r0 + d, d);

    // pipeline state: per stage the previous row's p (after boundaries), new vx, new vy
    double pb[K][4], un[K][4], vn[K][4];
    RowInfo info[K + 1];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) pb[s][c] = un[s][c] = vn[s][c] = 0.0;
#pragma unroll
    for (int s = 0; s <= K; ++s) info[s] = RowInfo{0u, 0, false, false};

    // owned cells of this lane: the 120 inner cells of the strip that lie inside the row
    const long long x0 = xs + 4 * lane;
    const bool lane_owned = lane >= 1 && lane <= 30 && x0 < nx;
    // shared-memory read pattern: two conflict-free 16-byte loads per field (lanes 4-7 of every
    // quarter-warp take the upper half first)
    const int swap = (lane >> 2) & 1;
    const int off_a = lane * 32 + swap * 16, off_b = lane * 32 + (swap ^ 1) * 16;

    int slot = 0;
    unsigned parity = 0;
    for (long long r = r0; r < r1; ++r) {
        mbar_wait(&bars[slot], parity);
        const unsigned char *src = ring + slot * kSlotBytes;
        double cur[3][4];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double2 va = *reinterpret_cast<const double2 *>(src + f * kStripCells * 8 + off_a);
            const double2 vb = *reinterpret_cast<const double2 *>(src + f * kStripCells * 8 + off_b);
            cur[f][0] = swap ? vb.x : va.x;
            cur[f][1] = swap ? vb.y : va.y;
            cur[f][2] = swap ? va.x : vb.x;
            cur[f][3] = swap ? va.y : vb.y;
        }
        const long long base = r * nx + xs;
        const unsigned idw = *reinterpret_cast<const unsigned *>(
            src + 3 * kStripCells * 8 + (int)(base & 15LL) + 4 * lane);
        __syncwarp();
        if (lane == 0 && r + kRingDepth < r1) issue(r + kRingDepth, slot);
        if (++slot == kRingDepth) { slot = 0; parity ^= 1; }

        // metadata of the new row (once per row, reused by all K stages)
#pragma unroll
        for (int s = K; s > 0; --s) info[s] = info[s - 1];
        {
            const unsigned first = __shfl_sync(0xffffffffu, idw, 0) & kIdMask;
            const bool uni = __all_sync(0xffffffffu, (idw & 0x3f3f3f3fu) == first * 0x01010101u);
            info[0].ids = idw;
            info[0].uniform = uni ? (int)first : -1;
            info[0].any_bound = __any_sync(0xffffffffu, idw & 0x40404040u);
            info[0].any_probe = __any_sync(0xffffffffu, idw & 0x80808080u);
        }

#pragma unroll
        for (int s = 0; s < K; ++s) {
            // stage s: cur = row q at level s  ->  cur = row q-1 at level s+1
            const long long q = r - s;
            const RowInfo &ri = info[s], &rp = info[s + 1];
            const long long cell0 = q * nx + x0;
            const long long sig = a.sig_index + s;
            double *__restrict__ record = t.ring + (a.ring_row + s) * t.n_slots;
            const bool owned_row = lane_owned && q >= ys && q < ye;

            // 1. boundaries and probes of p (row q)
            if (ri.any_bound) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((ri.ids >> (8 * c)) & kFlagBound)
                        cur[0][c] = apply_bounds(t.bound[0], t.signals, t.sig_steps, sig, cell0 + c,
                                                 cur[0][c]);
            }
            if (ri.any_probe && owned_row) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((ri.ids >> (8 * c)) & kFlagProbe)
                        write_probes(t.probe[0], record, cell0 + c, cur[0][c]);
            }

            // 2. new vx, vy of row q (backward differences of p), their boundaries and probes;
            // 3. new p of row q-1 (forward differences of the new vx, vy).
            // `coef` yields the material coefficient of a cell: gx(c) for row q, cell c-1 (c = 0 is
            // the left-hand lane's last cell); gyc/fyc(c) row q; gyp/fxp/fyp(c) row q-1 (fxp(4) is the
            // right-hand lane's first cell).
            double nu[4], nv[4], np[4];
            auto math = [&](const auto &coef) {
                const double p_left = shfl_up1(cur[0][3]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double pl = c ? cur[0][c - 1] : p_left;
                    nu[c] = sub(cur[1][c], diff2(coef.gx(c), pl, coef.gx(c + 1), cur[0][c]));
                    nv[c] = sub(cur[2][c], diff2(coef.gyp(c), pb[s][c], coef.gyc(c), cur[0][c]));
                }
                if (ri.any_bound) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if ((ri.ids >> (8 * c)) & kFlagBound) {
                            nu[c] = apply_bounds(t.bound[1], t.signals, t.sig_steps, sig, cell0 + c,
                                                 nu[c]);
                            nv[c] = apply_bounds(t.bound[2], t.signals, t.sig_steps, sig, cell0 + c,
                                                 nv[c]);
                        }
                }
                if (ri.any_probe && owned_row) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if ((ri.ids >> (8 * c)) & kFlagProbe) {
                            write_probes(t.probe[1], record, cell0 + c, nu[c]);
                            write_probes(t.probe[2], record, cell0 + c, nv[c]);
                        }
                }
                const double u_right = shfl_down1(un[s][0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double ur = c < 3 ? un[s][c + 1] : u_right;
                    const double divx = diff2(coef.fxp(c), un[s][c], coef.fxp(c + 1), ur);
                    const double divy = diff2(coef.fyp(c), vn[s][c], coef.fyc(c), nv[c]);
                    np[c] = sub(pb[s][c], add(divx, divy));
                }
            };
            if (ri.uniform >= 0 && ri.uniform == rp.uniform) {
                // all 256 cells of rows q-1 and q share one material: four coefficients in registers
                struct {
                    double g0, g1, f0, f1;
                    __device__ double gx(int) const { return g0; }
                    __device__ double gyc(int) const { return g1; }
                    __device__ double gyp(int) const { return g1; }
                    __device__ double fxp(int) const { return f0; }
                    __device__ double fyp(int) const { return f1; }
                    __device__ double fyc(int) const { return f1; }
                } coef{tabs[FDS_TAB_GX][ri.uniform], tabs[FDS_TAB_GY][ri.uniform],
                       tabs[FDS_TAB_FX][ri.uniform], tabs[FDS_TAB_FY][ri.uniform]};
                math(coef);
            } else {
                struct {
                    const double (*tabs)[kMaxMaterials];
                    unsigned cur_ids, prev_ids, left, right;
                    __device__ int mc(int c) const { return (cur_ids >> (8 * c)) & kIdMask; }
                    __device__ int mp(int c) const { return (prev_ids >> (8 * c)) & kIdMask; }
                    __device__ double gx(int c) const {
                        return tabs[FDS_TAB_GX][c ? mc(c - 1) : (int)((left >> 24) & kIdMask)];
                    }
                    __device__ double gyc(int c) const { return tabs[FDS_TAB_GY][mc(c)]; }
                    __device__ double gyp(int c) const { return tabs[FDS_TAB_GY][mp(c)]; }
                    __device__ double fxp(int c) const {
                        return tabs[FDS_TAB_FX][c < 4 ? mp(c) : (int)(right & kIdMask)];
                    }
                    __device__ double fyp(int c) const { return tabs[FDS_TAB_FY][mp(c)]; }
                    __device__ double fyc(int c) const { return tabs[FDS_TAB_FY][mc(c)]; }
                } coef{tabs, ri.ids, rp.ids,
                       __shfl_up_sync(0xffffffffu, ri.ids, 1),       // lane-1, row q
                       __shfl_down_sync(0xffffffffu, rp.ids, 1)};    // lane+1, row q-1
                math(coef);
            }

            // 4. hand row q-1 (level s+1) to the next stage, keep row q for the next iteration
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double keep_p = cur[0][c];
                cur[0][c] = np[c];
                cur[1][c] = un[s][c];
                cur[2][c] = vn[s][c];
                pb[s][c] = keep_p;
                un[s][c] = nu[c];
                vn[s][c] = nv[c];
            }
        }

        // row r-K at level K
        const long long orow = r - K;
        if (lane_owned && orow >= ys && orow < ye) {
            const long long o = orow * nx + x0;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                *reinterpret_cast<double2 *>(a.out[f] + o) = make_double2(cur[f][0], cur[f][1]);
                *reinterpret_cast<double2 *>(a.out[f] + o + 2) = make_double2(cur[f][2], cur[f][3]);
            }
        }
    }
}

}  // namespace fds
