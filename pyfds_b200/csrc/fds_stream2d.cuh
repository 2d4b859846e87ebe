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
#ifndef FDS_STREAM_MIN_CTAS
#define FDS_STREAM_MIN_CTAS 2
#endif
#ifndef FDS_STREAM_RING
#define FDS_STREAM_RING 6
#endif
constexpr int kStreamCtasPerSm = FDS_STREAM_MIN_CTAS;   // resident CTAs per SM (register budget)
constexpr int kRingDepth = FDS_STREAM_RING;   // rows in flight per warp
constexpr int kStreamWarps = 4;      // warps per CTA
constexpr int kMaxStreamSteps = 4;   // K
constexpr int kMapWindowBytes = kStripCells * 2 + 16;         // 128 map entries + alignment slack
constexpr int kSlotBytes = 3 * kStripCells * 8 + kMapWindowBytes + 16;   // p, vx, vy rows + map
constexpr int kScratchBytes = 32 * 64;                        // slow path: 8 doubles per lane
constexpr int kWarpRingBytes = kRingDepth * kSlotBytes + kScratchBytes + 64;   // + mbarriers

struct Stream2DArgs {
    const double *in[3];
    double *out[3];
    const map_t *map;      // material id + flags + classes, origin at local cell 0
    const double *tab;     // coefficient tables [FDS_TAB_COUNT][kMaxMaterials]
    const StepTables *tables;   // device copy, read on the slow path (boundaries, probes) only
    int *task_counter;     // dynamic task distribution (zero before the launch)
    const int *strip_order;   // strips sorted by expected cost, most expensive first
    long long nx;          // row length, multiple of 4
    long long row_begin;   // rows [row_begin, row_end) are produced
    long long row_end;
    int chunk_rows;        // rows per task
    int n_strips;
    int n_chunks;
    int n_tasks;
    long long sig_index;   // first step - sig_first_step
    long long ring_row;    // probe record of the first step
    int write_vector;      // thermal: store the flux components of the last step as well
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

// Slow path, taken only by rows that carry a boundary operation or a probe: applies the boundary
// operations of `n_comp` consecutive components (first_comp, first_comp + 1) to this lane's 4 cells and
// records probes. Values travel through a per-lane shared-memory scratch so that the call site stays a
// handful of instructions (the pipeline itself must fit the instruction cache).
__device__ __noinline__ void stream_slow_cells(const StepTables *__restrict__ tp, int first_comp,
                                               int n_comp, long long sig, long long record_row,
                                               long long cell0, unsigned long long ids, bool owned,
                                               double *scratch) {
    for (int k = 0; k < n_comp; ++k) {
        const int comp = first_comp + k;
        for (int c = 0; c < 4; ++c) {
            const unsigned f = (unsigned)(ids >> (16 * c));
            if (!(f & (kFlagBound | kFlagProbe))) continue;
            double v = scratch[4 * k + c];
            if (f & kFlagBound) {
                v = apply_bounds(tp->bound[comp], tp->rows, tp->signals, tp->sig_steps, sig,
                                 cell0 + c, v);
                scratch[4 * k + c] = v;
            }
            if ((f & kFlagProbe) && owned)
                write_probes(tp->probe[comp], tp->rows, tp->ring + record_row * tp->n_slots,
                             cell0 + c, v);
        }
    }
}

// Row metadata that travels through the pipeline with the row.
struct RowInfo {
    unsigned long long ids;   // the 4 map entries of this lane's cells
    int uniform;       // material id shared by all 128 cells of the row, or -1
    bool flagged;      // some cell of the row (any lane) needs the slow path (table lookup, probe)
    unsigned classed;  // bit c set: some cell of the row (any lane) carries an inline constant
                       // boundary operation on component c
};

// THERMAL: Thermal2D (pyfds/thermal.py:92-107) -- same pipeline with the temperature as the only
// state: the flux components are not read (q = -(A_q_t T) overwrites them, they are not accumulated)
// and only stored when the host asks for them after the last step of a call.
template <int K, bool THERMAL>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamCtasPerSm)
stream2d_kernel(Stream2DArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double tabs[4][kMaxMaterials];   // FDS_TAB_GX, GY, FX, FY
    __shared__ double cls_alpha[3][kMaxClasses], cls_value[3][kMaxClasses];

    for (int k = threadIdx.x; k < 4 * kMaxMaterials; k += blockDim.x) (&tabs[0][0])[k] = a.tab[k];
    for (int k = threadIdx.x; k < 3 * kMaxClasses; k += blockDim.x) {
        (&cls_alpha[0][0])[k] = (&a.tables->cls_alpha[0][0])[k];
        (&cls_value[0][0])[k] = (&a.tables->cls_value[0][0])[k];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nx = a.nx;
    unsigned char *ring = smem_raw + warp * kWarpRingBytes;
    double *scratch = reinterpret_cast<double *>(ring + kRingDepth * kSlotBytes) + lane * 8;
    unsigned long long *bars =
        reinterpret_cast<unsigned long long *>(ring + kRingDepth * kSlotBytes + kScratchBytes);

    // shared-memory read pattern: two conflict-free 16-byte loads per field (lanes 4-7 of every
    // quarter-warp take the upper half first)
    const int swap = (lane >> 2) & 1;
    const int off_a = lane * 32 + swap * 16, off_b = lane * 32 + (swap ^ 1) * 16;
    unsigned phase_bits = 0;   // parity of every ring slot (the barriers live across tasks)

    if (lane == 0) {
        for (int d = 0; d < kRingDepth; ++d) mbar_init(&bars[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    for (;;) {
        // ---- next task: (strip, chunk of rows), handed out dynamically ---------------------------
        int task = 0;
        if (lane == 0) task = atomicAdd(a.task_counter, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const int strip = __ldg(a.strip_order + task / a.n_chunks);
        const long long chunk = task % a.n_chunks;
        const long long ys = a.row_begin + chunk * a.chunk_rows;
        const long long ye = min(ys + (long long)a.chunk_rows, a.row_end);
        const long long xs = (long long)strip * kStripStride - kStripHalo;   // column of lane 0
        const long long r0 = ys - K, r1 = ye + K;                            // rows streamed in

        auto issue = [&](long long r, int slot) {
            const long long base = r * nx + xs;         // flat cell index of the strip start
            const long long base8 = base & ~7LL;        // map entries: 16-byte aligned window
            unsigned char *dst = ring + slot * kSlotBytes;
            mbar_expect_tx(&bars[slot], (THERMAL ? 1 : 3) * kStripCells * 8 + kMapWindowBytes);
            bulk_load(dst, a.in[0] + base, kStripCells * 8, &bars[slot]);
            if (!THERMAL) {
                bulk_load(dst + kStripCells * 8, a.in[1] + base, kStripCells * 8, &bars[slot]);
                bulk_load(dst + 2 * kStripCells * 8, a.in[2] + base, kStripCells * 8, &bars[slot]);
            }
            bulk_load(dst + 3 * kStripCells * 8, a.map + base8, kMapWindowBytes, &bars[slot]);
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
        for (int s = 0; s <= K; ++s) info[s] = RowInfo{0ull, 0, false, 0u};

        // owned cells of this lane: the 120 inner cells of the strip that lie inside the row
        const long long x0 = xs + 4 * lane;
        const bool lane_owned = lane >= 1 && lane <= 30 && x0 < nx;
        long long cell_r = r0 * nx + x0;    // flat index of this lane's first cell in row r
        const int map_step = (int)(nx & 7LL);
        int map_off = (int)((r0 * nx + xs) & 7LL);   // (r * nx + xs) & 7: entry offset in the window

        int slot = 0;
        for (long long r = r0; r < r1; ++r) {
            mbar_wait(&bars[slot], (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            const unsigned char *src = ring + slot * kSlotBytes;
            double cur[3][4];
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                if (THERMAL && f > 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) cur[f][c] = 0.0;
                    continue;
                }
                const double2 va =
                    *reinterpret_cast<const double2 *>(src + f * kStripCells * 8 + off_a);
                const double2 vb =
                    *reinterpret_cast<const double2 *>(src + f * kStripCells * 8 + off_b);
                cur[f][0] = swap ? vb.x : va.x;
                cur[f][1] = swap ? vb.y : va.y;
                cur[f][2] = swap ? va.x : vb.x;
                cur[f][3] = swap ? va.y : vb.y;
            }
            const unsigned long long idw = *reinterpret_cast<const unsigned long long *>(
                src + 3 * kStripCells * 8 + 2 * map_off + 8 * lane);
            map_off = (map_off + map_step) & 7;
            __syncwarp();
            if (lane == 0 && r + kRingDepth < r1) issue(r + kRingDepth, slot);
            if (++slot == kRingDepth) slot = 0;

            // metadata of the new row (once per row, reused by all K stages)
#pragma unroll
            for (int s = K; s > 0; --s) info[s] = info[s - 1];
            {
                const unsigned long long first = __shfl_sync(0xffffffffu, idw, 0) & kIdMask;
                const bool uni = __all_sync(0xffffffffu, (idw & 0x001f001f001f001full) ==
                                                             first * 0x0001000100010001ull);
                info[0].ids = idw;
                info[0].uniform = uni ? (int)first : -1;
                info[0].flagged = __any_sync(0xffffffffu, (idw & 0x0060006000600060ull) != 0);
                info[0].classed = 0u;
                if (__any_sync(0xffffffffu, (idw & 0xff80ff80ff80ff80ull) != 0)) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const unsigned long long m = 0x0007000700070007ull << class_shift(c);
                        if (__any_sync(0xffffffffu, (idw & m) != 0)) info[0].classed |= 1u << c;
                    }
                }
            }

#pragma unroll
            for (int s = 0; s < K; ++s) {
                // stage s: cur = row q = r - s at level s  ->  cur = row q-1 at level s+1
                const RowInfo &ri = info[s], &rp = info[s + 1];

                // 1. boundaries and probes of p (row q)
                if (ri.classed & 1u) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        cur[0][c] = apply_class(cls_alpha, cls_value, 0,
                                                (unsigned)(ri.ids >> (16 * c)), cur[0][c]);
                }
                if (ri.flagged) {
                    const long long q = r - s;
#pragma unroll
                    for (int c = 0; c < 4; ++c) scratch[c] = cur[0][c];
                    stream_slow_cells(a.tables, 0, 1, a.sig_index + s, a.ring_row + s,
                                      cell_r - s * nx, ri.ids,
                                      lane_owned && q >= ys && q < ye, scratch);
#pragma unroll
                    for (int c = 0; c < 4; ++c) cur[0][c] = scratch[c];
                }

                // 2. new vx, vy of row q (backward differences of p), their boundaries and probes;
                // 3. new p of row q-1 (forward differences of the new vx, vy).
                // `coef` yields the material coefficient of a cell: gx(c) for row q, cell c-1 (c = 0
                // is the left-hand lane's last cell); gyc/fyc(c) row q; gyp/fxp/fyp(c) row q-1
                // (fxp(4) is the right-hand lane's first cell).
                double nu[4], nv[4], np[4];
                auto math = [&](const auto &coef) {
                    const double p_left = shfl_up1(cur[0][3]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double pl = c ? cur[0][c - 1] : p_left;
                        const double dx_ = diff2(coef.gx(c), pl, coef.gx(c + 1), cur[0][c]);
                        const double dy_ = diff2(coef.gyp(c), pb[s][c], coef.gyc(c), cur[0][c]);
                        nu[c] = THERMAL ? -dx_ : sub(cur[1][c], dx_);
                        nv[c] = THERMAL ? -dy_ : sub(cur[2][c], dy_);
                    }
                    if (ri.classed & 2u) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            nu[c] = apply_class(cls_alpha, cls_value, 1,
                                                (unsigned)(ri.ids >> (16 * c)), nu[c]);
                    }
                    if (ri.classed & 4u) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            nv[c] = apply_class(cls_alpha, cls_value, 2,
                                                (unsigned)(ri.ids >> (16 * c)), nv[c]);
                    }
                    if (ri.flagged) {
                        const long long q = r - s;
#pragma unroll
                        for (int c = 0; c < 4; ++c) { scratch[c] = nu[c]; scratch[4 + c] = nv[c]; }
                        stream_slow_cells(a.tables, 1, 2, a.sig_index + s, a.ring_row + s,
                                          cell_r - s * nx, ri.ids,
                                          lane_owned && q >= ys && q < ye, scratch);
#pragma unroll
                        for (int c = 0; c < 4; ++c) { nu[c] = scratch[c]; nv[c] = scratch[4 + c]; }
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
                    // all 256 cells of rows q-1 and q share one material: 4 coefficients in registers
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
                        unsigned long long cur_ids, prev_ids, left, right;
                        __device__ int mc(int c) const { return (int)(cur_ids >> (16 * c)) & kIdMask; }
                        __device__ int mp(int c) const { return (int)(prev_ids >> (16 * c)) & kIdMask; }
                        __device__ double gx(int c) const {
                            return tabs[FDS_TAB_GX][c ? mc(c - 1) : (int)((left >> 48) & kIdMask)];
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
                const long long o = cell_r - K * nx;
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (THERMAL && f > 0 && !a.write_vector) continue;
                    *reinterpret_cast<double2 *>(a.out[f] + o) = make_double2(cur[f][0], cur[f][1]);
                    *reinterpret_cast<double2 *>(a.out[f] + o + 2) =
                        make_double2(cur[f][2], cur[f][3]);
                }
            }
            cell_r += nx;
        }
    }
}

}  // namespace fds
