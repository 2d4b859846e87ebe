// Streaming multi-step kernel for the lossless Acoustic2D leapfrog (pyfds/acoustics.py:111-128), for
// Thermal2D (pyfds/thermal.py:92-107) and for their axisymmetric counterparts (lossless Acoustic3DAxi,
// pyfds/acoustics.py:205-225; Thermal3DAxi, pyfds/thermal.py:160-176): K time steps per launch with
// ONE read and ONE write of the state (temporal blocking).
//
// Work decomposition -- every warp is autonomous, there is no block-level synchronisation in the loop:
//   * a warp owns an x-strip of 56 cells and a chunk of rows; it streams the strip (64 cells wide:
//     4 halo cells either side) row by row through a K-stage time pipeline held in REGISTERS.
//     Stage t receives row q at time level t and emits row q-1 at level t+1, so after K stages the
//     warp stores row r-K at level K while row r is being read. Per stage only three row-fragments
//     stay live (p after boundaries, new vx, new vy of the previous row): 12 registers per lane.
//   * each lane holds 2 consecutive cells (one 16-byte word); the x-neighbours that live in the
//     adjacent lane travel by warp shuffle (p to the right-hand lane for the backward difference, new
//     vx to the left-hand lane for the forward difference). The two outermost lanes either side
//     compute garbage that never reaches the 56 owned cells: the dependency cone grows one cell per
//     step and K <= 4. Two cells per lane keep the kernel at 128 registers, i.e. four CTAs = 16 warps
//     per SM (a 4-cell variant with 2 CTAs per SM ran the FP64 pipe at 35 %, this one at 52 %).
//   * rows are fetched in pairs by the bulk-copy engine (TMA, cp.async.bulk + mbarrier) into a
//     per-warp ring of shared-memory slots, two pairs ahead of the arithmetic, so DRAM latency is
//     hidden without spending registers on loads in flight. Strips are addressed by flat cell index,
//     which gives the reference's row wrap (pyfds/fields.py:290-297) and the zero padding at the grid
//     ends for free.
//   * STEADY rows -- the lane's map word repeats row after row, nothing needs a table lookup -- run
//     through a branch-free body that consumes two rows per iteration with the pipeline state
//     alternating between two register sets (no register moves between iterations). Five
//     instantiations: one material (coefficients warp-uniform) with constant boundary operations on
//     no component or on exactly one (applied to all cells as v = alpha*v + value, the identity
//     being alpha = 1, value = -0.0), and several materials without operations (per-lane
//     coefficients). Everything else -- rows around a change of the map, sources with signals,
//     probes -- takes the general row iteration: one rolled copy of the stage code with per-cell
//     coefficient lookups, inline classes and the table slow path.
//   * tasks (strip, rows) come from a host-built table balanced with a device census of the rows that
//     are not steady (fds_abi.cu: build_stream_plan) and are handed out dynamically.
//
// Arithmetic is the same sequence of __dmul_rn/__dadd_rn as the one-step kernel, so results are
// bitwise identical to it and to the reference (tests/test_gpu_parity.py).
#pragma once

#include <type_traits>

#include "fds_common.cuh"

namespace fds {

#ifndef FDS_STREAM_WARPS
#define FDS_STREAM_WARPS 4
#endif
constexpr int kStreamWarps = FDS_STREAM_WARPS;   // warps per CTA
constexpr int kMaxStreamSteps = 4;   // K

// ---- ordering between tasks: overlapped sweeps and the in-kernel halo exchange ------------------------
// A launch (one "sweep" of K steps over the slab) no longer has to wait for the whole previous sweep:
//   * launched with programmatic stream serialisation, its CTAs become resident as soon as CTAs of the
//     previous sweep exit, and every task waits only for the tasks of the previous sweep whose output it
//     reads (and which read what it is about to overwrite -- the same set): `dep_idx[dep_ptr[t] ..
//     dep_ptr[t+1])`, host-built from the task table, checked against `done[]`, which every task
//     releases with the sweep number when its rows are stored. The tail of one sweep (the slowest SMs,
//     the last tasks) is filled with the head of the next one.
//   * multi-GPU: tasks that read halo rows wait for the flag the neighbour slab releases when its edge
//     rows of the previous sweep have landed in this slab's halo rows; tasks that own edge rows copy
//     them straight into the neighbour's halo rows (peer memory over NVLink) when they finish, and the
//     last one of a side releases the neighbour's flag. Edge tasks are handed out first, so the rows
//     travel while the interior of the slab is still being computed.
// Task flags (int4::w of the task table):
constexpr int kTaskWaitLo = 1, kTaskWaitHi = 2;    // reads halo rows of the lower / upper neighbour
constexpr int kTaskPushLo = 4, kTaskPushHi = 8;    // owns rows the lower / upper neighbour needs

struct TaskSync {
    unsigned seq;               // number of this sweep (> 0, counted per context)
    int wait_deps;              // 1: the previous sweep used this task table and may still be running
    const int *dep_ptr;         // CSR over the tasks of the table; nullptr: no dependency lists
    const int *dep_idx;
    unsigned *done;             // [n_tasks] last sweep each task finished; nullptr: not recorded
    // in-kernel halo exchange (all null / 0 on a single GPU)
    unsigned halo_wait_seq;     // flags must have reached this sweep number (0: nothing to wait for)
    int halo_push;              // 1: edge tasks push their rows and release the neighbours' flags
    const unsigned *flag_in[2]; // this slab's flags, written by the lower / upper neighbour
    unsigned *flag_out[2];      // the neighbours' flags (peer memory)
    double *peer_dst[2][3];     // [side][component]: where row 0 of the side's edge band goes
    unsigned *edge_done;        // [side][sweep parity] counters of finished edge tasks
    unsigned *error;            // set to 1 when a wait gave up
    int n_edge_tasks[2];
    int halo_rows;              // rows per edge band
    int ncomp;                  // components that travel (thermal: 1)
    long long rows;             // owned rows of the slab
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spins until *flag >= want (wrapping comparison); gives up after 20 s and raises *error.
template <bool SYSTEM>
__device__ __noinline__ void spin_until(const unsigned *flag, unsigned want, unsigned *error) {
    unsigned seen;
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        if (SYSTEM)
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        else
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if ((int)(seen - want) >= 0) return;
        if (spins > 64) __nanosleep(64);
        if ((spins & 1023u) == 1023u) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) {
                if (error) atomicExch(error, 1u);
                return;
            }
        }
    }
}

// Before a task touches memory: wait for what it reads to be there and for what it overwrites to have
// been read. Called by the whole warp; what follows (bulk loads issued by lane 0 through the async
// proxy) is ordered behind the acquires by the proxy fence.
__device__ __forceinline__ void task_acquire(const TaskSync &y, int task, int flags, int lane) {
    bool waited = false;
    if (y.halo_wait_seq && (flags & (kTaskWaitLo | kTaskWaitHi))) {
        if (lane < 2 && (flags & (kTaskWaitLo << lane)) && y.flag_in[lane])
            spin_until<true>(y.flag_in[lane], y.halo_wait_seq, y.error);
        waited = true;
    }
    if (y.wait_deps && y.dep_ptr) {
        const int end = __ldg(y.dep_ptr + task + 1);
        for (int k = __ldg(y.dep_ptr + task) + lane; k < end; k += 32)
            spin_until<false>(y.done + __ldg(y.dep_idx + k), y.seq - 1u, y.error);
        waited = true;
    }
    if (waited) {
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");
    }
}

// After the last row of a task has been stored (by all lanes of the warp).
__device__ __noinline__ void task_release(const TaskSync &y, double *const *out, long long nx,
                                          int task, int strip_first_col, int ys, int ye, int flags,
                                          int lane) {
    __syncwarp();
    if (y.halo_push && (flags & (kTaskPushLo | kTaskPushHi))) {
        // 56 owned cells of the strip = 28 16-byte words per row and component
        const long long x = strip_first_col + 2 * lane;
        const bool mine = lane < 28 && x < nx;
        for (int side = 0; side < 2; ++side) {
            if (!(flags & (kTaskPushLo << side)) || !y.flag_out[side]) continue;
            const long long band0 = side ? y.rows - y.halo_rows : 0;   // first row of the edge band
            const long long lo = ys > band0 ? ys : band0;
            const long long hi = ye < band0 + y.halo_rows ? ye : band0 + y.halo_rows;
            if (mine)
                for (long long row = lo; row < hi; ++row)
                    for (int c = 0; c < y.ncomp; ++c) {
                        const double2 v =
                            __ldcg(reinterpret_cast<const double2 *>(out[c] + row * nx + x));
                        *reinterpret_cast<double2 *>(y.peer_dst[side][c] + (row - band0) * nx + x) = v;
                    }
            __syncwarp();
            if (lane == 0) {
                __threadfence_system();
                unsigned *counter = y.edge_done + 2 * side + (y.seq & 1u);
                const unsigned before = atomicAdd(counter, 1u);
                if (before == (unsigned)y.n_edge_tasks[side] - 1u) {
                    atomicExch(counter, 0u);
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(y.flag_out[side]),
                                 "r"(y.seq)
                                 : "memory");
                }
            }
        }
    }
    if (y.done && lane == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(y.done + task), "r"(y.seq)
                     : "memory");
    }
}

struct Stream2DArgs {
    const double *in[3];
    double *out[3];
    const map_t *map;      // material id + flags + classes, origin at local cell 0
    const double *tab;     // coefficient tables [FDS_TAB_COUNT][kMaxMaterials]
    const StepTables *tables;   // device copy, read on the slow path (boundaries, probes) only
    int *task_counter;     // dynamic task distribution (zero before the launch)
    const int4 *tasks;     // (strip, first row, end row, -) per task, most expensive first
    long long nx;          // row length, multiple of 4
    long long row_begin;   // rows [row_begin, row_end) are produced
    long long row_end;
    int n_strips;
    int n_tasks;
    long long sig_index;   // first step - sig_first_step
    long long ring_row;    // probe record of the first step
    int write_vector;      // thermal: store the flux components of the last step as well
    // optional counters (nullptr unless FDS_STREAM_STATS is set): [0..4] entries into the branch-free
    // body by variant (no operation, operations on component 0 / 1 / 2, several materials),
    // [5] rows through the general row iteration, [6] rows streamed in total
    unsigned long long *stats;
    // axisymmetric models (Acoustic3DAxi lossless, Thermal3DAxi): per-(material, column) and per-column
    // values, as in StepTables
    const double *ctab;    // [FDS_CTAB_COUNT][n_mat1][nx]
    const double *cvec;    // [FDS_CVEC_COUNT][nx]
    int n_mat1;
    TaskSync sync;         // ordering between the tasks of consecutive sweeps and of neighbour slabs
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

// =====================================================================================================
// stream2d_kernel: lossless Acoustic2D / Thermal2D, K steps per launch
// =====================================================================================================
// Geometry of this kernel: a lane holds kS2LaneCells = 2 consecutive cells, so a warp streams a strip
// of 64 cells (two halo lanes = 4 cells either side, 56 owned). Two cells per lane keep the K = 4
// pipeline state at 48 registers and the whole kernel under 128, i.e. FOUR resident CTAs (16 warps)
// per SM: the FP64 pipe is fed by four warps per scheduler instead of two (measured: the 4-cell
// variant ran at 35 % FP64 pipe utilisation, bound by issue stalls of its two warps).
constexpr int kS2LaneCells = 2;
constexpr int kS2StripCells = 32 * kS2LaneCells;                   // 64
constexpr int kS2StripHalo = 4;                                    // cells; K <= 4
constexpr int kS2HaloLanes = kS2StripHalo / kS2LaneCells;          // 2
constexpr int kS2StripStride = kS2StripCells - 2 * kS2StripHalo;   // 56 owned cells per strip
#ifndef FDS_S2_CTAS
#define FDS_S2_CTAS (12 / FDS_STREAM_WARPS)   // 12 warps per SM at 168 registers
#endif
#ifndef FDS_S2_RING
#define FDS_S2_RING 6
#endif
constexpr int kS2CtasPerSm = FDS_S2_CTAS;   // 3 CTAs x 4 warps at 168 registers: no spills
constexpr int kS2RingDepth = FDS_S2_RING;   // rows in flight per warp (whole pairs)
constexpr int kS2MapWindowBytes = kS2StripCells * 2 + 16;
constexpr int kS2FieldBytes = kS2StripCells * 8;                   // 512
constexpr int kS2SlotBytes = 3 * kS2FieldBytes + kS2MapWindowBytes + 16;
constexpr int kS2ScratchBytes = 32 * 2 * kS2LaneCells * 8;         // slow path: 2 components per lane
constexpr int kS2WarpRingBytes = kS2RingDepth * kS2SlotBytes + kS2ScratchBytes + 64;   // + mbarriers, task stash
static_assert(8 * (kS2RingDepth / 2) + 8 <= 64, "mbarriers and the task stash share 64 bytes");
static_assert(kS2SlotBytes % 16 == 0, "bulk copies need 16-byte aligned slots");
static_assert(kS2LaneCells == 2, "the 32-bit lane map word and the double2 row accesses assume 2");

// Row metadata, packed (the kernel keeps K + 3 of them in registers).
struct RowTag {
    unsigned ids;    // the 2 map entries of this lane's cells
    unsigned bits;   // 0-4 material of the row's first cell, 5 row is not of one material,
                     // 6 flagged, 7-9 classed (bit 7 + c: component c)
    __device__ int uniform() const { return (bits & 32u) ? -1 : (int)(bits & 31u); }
    __device__ bool flagged() const { return (bits & 64u) != 0; }
    __device__ unsigned classed() const { return bits >> 7; }
    // plain: one material, no boundary operation, no probe
    __device__ bool plain() const { return bits < 32u; }
    // steady: no table lookup, no probe, and either constant operations on at most ONE component of
    // a row of one material, or several materials without any operation -- such rows take the
    // branch-free body as long as every lane sees the same map word row after row (a wall, a constant
    // source or a material interface along y; plain rows are the special case without any of it)
    __device__ bool steady() const {
        const unsigned cls = bits >> 7;
        return (bits & 64u) == 0u && ((bits & 32u) ? cls == 0u : (cls & (cls - 1u)) == 0u);
    }
};

// Metadata of the row in ring slot `src` (once per row, reused by all K stages). Lanes whose cells lie
// beyond the dependency cone of the owned cells (`relevant` false: past the row end plus the strip
// halo) do not vote: what they compute is never used.
__device__ __forceinline__ RowTag s2_row_meta(const unsigned char *src, int map_off, int lane,
                                              bool relevant) {
    const unsigned raw = *reinterpret_cast<const unsigned *>(src + 3 * kS2FieldBytes + 2 * map_off +
                                                             4 * lane);
    const unsigned first = __shfl_sync(0xffffffffu, raw, 0) & kIdMask;
    // irrelevant lanes take the row's first material and no flags
    const unsigned idw = relevant ? raw : first * 0x00010001u;
    RowTag m;
    const bool uni = __all_sync(0xffffffffu, (idw & 0x001f001fu) == first * 0x00010001u);
    m.ids = idw;
    m.bits = first | (uni ? 0u : 32u);
    if (__any_sync(0xffffffffu, (idw & 0xffe0ffe0u) != 0)) {
        if (__any_sync(0xffffffffu, (idw & 0x00600060u) != 0)) m.bits |= 64u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned mask = 0x00070007u << class_shift(c);
            if (__any_sync(0xffffffffu, (idw & mask) != 0)) m.bits |= 128u << c;
        }
    }
    return m;
}

// Class tables of one launch in shared memory: alpha by class, value by stage and class (stage s works
// on step sig_index + s, so a class that stands for a signal takes that sample).
template <int K>
__device__ __forceinline__ void stage_class_tables(const Stream2DArgs &a,
                                                   double (*cls_alpha)[kMaxClasses],
                                                   double (*cls_value)[3][kMaxClasses]) {
    const StepTables *__restrict__ t = a.tables;
    for (int k = threadIdx.x; k < 3 * kMaxClasses; k += blockDim.x) {
        (&cls_alpha[0][0])[k] = (&t->cls_alpha[0][0])[k];
        const int sidx = (&t->cls_signal[0][0])[k];
#pragma unroll
        for (int s = 0; s < K; ++s)
            (&cls_value[s][0][0])[k] =
                sidx >= 0 ? t->signals[(long long)sidx * t->sig_steps + a.sig_index + s]
                          : (&t->cls_value[0][0])[k];
    }
}

// Slow path, taken only by rows that carry a boundary operation or a probe: applies the boundary
// operations of `n_comp` consecutive components (first_comp, first_comp + 1) to this lane's 2 cells and
// records probes. Values travel through a per-lane shared-memory scratch so that the call site stays a
// handful of instructions (the pipeline itself must fit the instruction cache).
__device__ __noinline__ void s2_slow_cells(const StepTables *__restrict__ tp, int first_comp,
                                           int n_comp, long long sig, long long record_row,
                                           long long cell0, unsigned ids, bool owned,
                                           double *scratch) {
    for (int k = 0; k < n_comp; ++k) {
        const int comp = first_comp + k;
        for (int c = 0; c < kS2LaneCells; ++c) {
            const unsigned f = ids >> (16 * c);
            if (!(f & (kFlagBound | kFlagProbe))) continue;
            double v = scratch[kS2LaneCells * k + c];
            if (f & kFlagBound) {
                v = apply_bounds(tp->bound[comp], tp->rows, tp->signals, tp->sig_steps, sig,
                                 cell0 + c, v);
                scratch[kS2LaneCells * k + c] = v;
            }
            if ((f & kFlagProbe) && owned)
                write_probes(tp->probe[comp], tp->rows, tp->ring + record_row * tp->n_slots,
                             cell0 + c, v);
        }
    }
}

// One stage of the time pipeline on a STEADY row pair (rows q-1 and q of one material, no table
// lookup, no probe): straight-line arithmetic, no branch. `cur` = row q at level s on entry and
// row q-1 at level s+1 on exit; P/U/V = p (after boundaries), new vx, new vy of row q-1 (read),
// Pn/Un/Vn = the same of row q (written: the state of the next row iteration).
// CC >= 0: component CC carries constant operations v = alpha * v + value on some cells of the
// strip; ca/cv hold them for this lane's cells, cells without one get alpha = 1, value = -0.0, which
// returns v bit for bit (1 * v is exact and v + (-0.0) = v for every v, signed zeros included).
// gx[c] = x-gradient coefficient of cell c-1 (gx[0]: the left-hand lane's last cell), gy[c], fy[c] =
// y coefficients of cell c, fx[c] = x-divergence coefficient of cell c (fx[C]: the right-hand lane's
// first cell): the same for rows q-1 and q, because steady rows repeat their map words.
// AXI (pyfds/acoustics.py:205-225 without losses, pyfds/thermal.py:160-176): fx[] holds the
// per-column a_p_vx / r (a_t_qx / r) values, the x-divergence is applied to vx * r (rv[c] = r of cell c,
// rv[C]: the right-hand lane's first cell), and the lossless vx update carries the reference's
// (0 * vx) / r^2 term as + 0 * vx.
template <bool THERMAL, int CC, bool AXI = false>
__device__ __forceinline__ void steady_stage(double (&cur)[3][kS2LaneCells],
                                             const double (&P)[kS2LaneCells],
                                             const double (&U)[kS2LaneCells],
                                             const double (&V)[kS2LaneCells],
                                             double (&Pn)[kS2LaneCells], double (&Un)[kS2LaneCells],
                                             double (&Vn)[kS2LaneCells],
                                             const double (&gx)[kS2LaneCells + 1],
                                             const double (&gy)[kS2LaneCells],
                                             const double (&fx)[kS2LaneCells + 1],
                                             const double (&fy)[kS2LaneCells],
                                             const double (&ca)[kS2LaneCells],
                                             const double (&cv)[kS2LaneCells],
                                             const double (&rv)[kS2LaneCells + 1]) {
    constexpr int C = kS2LaneCells;
    if (CC == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) cur[0][c] = add(mul(ca[c], cur[0][c]), cv[c]);
    }
    const double p_left = shfl_up1(cur[0][C - 1]);
    double np[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double pl = c ? cur[0][c - 1] : p_left;
        const double dx_ = diff2(gx[c], pl, gx[c + 1], cur[0][c]);
        const double dy_ = diff2(gy[c], P[c], gy[c], cur[0][c]);
        Un[c] = THERMAL ? -dx_
                : AXI   ? sub(cur[1][c], add(dx_, mul(0.0, cur[1][c])))
                        : sub(cur[1][c], dx_);
        Vn[c] = THERMAL ? -dy_ : sub(cur[2][c], dy_);
        if (CC == 1) Un[c] = add(mul(ca[c], Un[c]), cv[c]);
        if (CC == 2) Vn[c] = add(mul(ca[c], Vn[c]), cv[c]);
    }
    const double u_right = shfl_down1(U[0]);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        double ul = U[c], ur = c < C - 1 ? U[c + 1] : u_right;
        if (AXI) {
            ul = mul(ul, rv[c]);
            ur = mul(ur, rv[c + 1]);
        }
        const double divx = diff2(fx[c], ul, fx[c + 1], ur);
        const double divy = diff2(fy[c], V[c], fy[c], Vn[c]);
        np[c] = sub(P[c], add(divx, divy));
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        Pn[c] = cur[0][c];
        cur[0][c] = np[c];
        cur[1][c] = U[c];
        cur[2][c] = V[c];
    }
}

// THERMAL: Thermal2D (pyfds/thermal.py:92-107) -- same pipeline with the temperature as the only
// state: the flux components are not read (q = -(A_q_t T) overwrites them, they are not accumulated)
// and only stored when the host asks for them after the last step of a call.
//
// Row metadata is fetched two rows ahead of the arithmetic. While the next two rows and the K+1 rows
// before them are PLAIN (one material, no boundary operation, no probe -- almost everywhere) rows
// are consumed in pairs by a branch-free body of 2 x K stages in which the pipeline state alternates
// between two register sets, so that no register is moved between row iterations; any other row
// takes the general row iteration.
// STATS: count in a.stats what the rows went through (tests only: a separate instantiation, so that
// the counters cannot disturb the register allocation of the production kernel).
template <int K, bool THERMAL, bool STATS = false, bool AXI = false>
__global__ void __launch_bounds__(kStreamWarps * 32, kS2CtasPerSm)
stream2d_kernel(Stream2DArgs a) {
    constexpr int C = kS2LaneCells;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double tabs[4][kMaxMaterials];   // FDS_TAB_GX, GY, FX, FY
    // class operations; the value by stage: a class may stand for a signal (one sample per step)
    __shared__ double cls_alpha[3][kMaxClasses], cls_value[K][3][kMaxClasses];

    for (int k = threadIdx.x; k < 4 * kMaxMaterials; k += blockDim.x) (&tabs[0][0])[k] = a.tab[k];
    stage_class_tables<K>(a, cls_alpha, cls_value);
    __syncthreads();
    // the next sweep may move in as soon as CTAs of this one leave (its tasks order themselves
    // against ours through a.sync); a no-op unless it was launched with programmatic serialisation
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nx = a.nx;
    unsigned char *ring = smem_raw + warp * kS2WarpRingBytes;
    double *scratch =
        reinterpret_cast<double *>(ring + kS2RingDepth * kS2SlotBytes) + lane * 2 * C;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        ring + kS2RingDepth * kS2SlotBytes + kS2ScratchBytes);
    // task number and flags wait here for the end of the task instead of occupying registers
    int *stash = reinterpret_cast<int *>(bars + kS2RingDepth / 2);
    // rows travel in pairs: ring slots 2j and 2j+1 share mbarrier j (one wait, one refill per pair)
    static_assert(kS2RingDepth % 2 == 0, "the ring holds whole row pairs");
    unsigned phase_bits = 0;   // parity of every pair barrier (the barriers live across tasks)

    if (lane == 0) {
        for (int d = 0; d < kS2RingDepth / 2; ++d) mbar_init(&bars[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    for (;;) {
        // ---- next task: (strip, rows), handed out dynamically, most expensive first --------------
        int task = 0;
        if (lane == 0) task = atomicAdd(a.task_counter, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const int4 tk = __ldg(a.tasks + task);
        const int ys = tk.y, ye = tk.z;
        const long long xs = (long long)tk.x * kS2StripStride - kS2StripHalo;   // column of lane 0
        const int r0 = ys - K, r1 = ye + K;                                     // rows streamed in
        if (lane == 0) {
            stash[0] = task;
            stash[1] = tk.w;
        }
        task_acquire(a.sync, task, tk.w, lane);

        // rows r, r+1 (r - r0 even) into ring slots `slot`, `slot` + 1 (slot even), one barrier
        auto issue_pair = [&](int r, int slot) {
            const int n_rows = r + 1 < r1 ? 2 : 1;
            void *bar = &bars[slot >> 1];
            mbar_expect_tx(bar, n_rows * ((THERMAL ? 1 : 3) * kS2FieldBytes + kS2MapWindowBytes));
            long long base = (long long)r * nx + xs;   // flat cell index of the strip start
            unsigned char *dst = ring + slot * kS2SlotBytes;
            for (int h = 0; h < n_rows; ++h) {
                bulk_load(dst, a.in[0] + base, kS2FieldBytes, bar);
                if (!THERMAL) {
                    bulk_load(dst + kS2FieldBytes, a.in[1] + base, kS2FieldBytes, bar);
                    bulk_load(dst + 2 * kS2FieldBytes, a.in[2] + base, kS2FieldBytes, bar);
                }
                // map entries: 16-byte aligned window
                bulk_load(dst + 3 * kS2FieldBytes, a.map + (base & ~7LL), kS2MapWindowBytes, bar);
                base += nx;
                dst += kS2SlotBytes;
            }
        };
        if (lane == 0)
            for (int d = 0; d < kS2RingDepth && r0 + d < r1; d += 2) issue_pair(r0 + d, d);
        if (STATS && lane == 0) atomicAdd(a.stats + 6, (unsigned long long)(r1 - r0));

        // pipeline state: per stage the previous row's p (after boundaries), new vx, new vy
        double pb[K][C], un[K][C], vn[K][C];
        RowTag info[K + 1];
#pragma unroll
        for (int s = 0; s < K; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) pb[s][c] = un[s][c] = vn[s][c] = 0.0;
#pragma unroll
        for (int s = 0; s <= K; ++s) info[s] = RowTag{0u, 0u};
        int run = 0;

        // owned cells of this lane: the 56 inner cells of the strip that lie inside the row
        const long long x0 = xs + C * lane;
        const bool lane_owned = lane >= kS2HaloLanes && lane < 32 - kS2HaloLanes && x0 < nx;
        const bool lane_relevant = x0 < nx + kS2StripHalo;
        long long cell_r = (long long)r0 * nx + x0;   // flat index of this lane's first cell in row r
        const int map_step = (int)(nx & 7LL);
        int map_off = (int)(((long long)r0 * nx + xs) & 7LL);   // entry offset in the map window

        // axisymmetric: column of cell c (0 .. C) of this lane, wrapped into the row like the flat
        // index; a_p_vx / r of a material and column; r of a column
        auto col_of = [&](int c) {
            long long col = x0 + c;
            if (col < 0) col += nx;
            if (col >= nx) col -= nx;
            if (col >= nx) col -= nx;
            return col;
        };
        auto fx_col = [&](unsigned m, int c) {
            return __ldg(a.ctab + ((long long)FDS_CTAB_FX * a.n_mat1 + (m & kIdMask)) * nx + col_of(c));
        };
        auto r_col = [&](int c) { return __ldg(a.cvec + FDS_CVEC_R * nx + col_of(c)); };

        // a lane's 2 cells are one 16-byte word: rows are read and written with 128-bit accesses,
        // lane after lane, free of bank conflicts
        auto load_row = [&](double (&cur)[3][C], const unsigned char *src) {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                if (THERMAL && f > 0) {
                    cur[f][0] = cur[f][1] = 0.0;
                    continue;
                }
                const double2 v =
                    *reinterpret_cast<const double2 *>(src + f * kS2FieldBytes + lane * 16);
                cur[f][0] = v.x;
                cur[f][1] = v.y;
            }
        };
        // row `orow` at level K, held in `cur`; `o` = flat index of this lane's first cell in it
        auto store_row = [&](const double (&cur)[3][C], int orow, long long o) {
            if (lane_owned && orow >= ys && orow < ye) {
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (THERMAL && f > 0 && !a.write_vector) continue;
                    // streaming store: the rows are not read again by this launch and should not
                    // evict what little the L1 holds (spilled loop counters); measured +2 %
                    __stcs(reinterpret_cast<double2 *>(a.out[f] + o),
                           make_double2(cur[f][0], cur[f][1]));
                }
            }
        };

        // metadata is fetched two rows ahead of the arithmetic: m0 = row r, m1 = row r+1
        int fetch_row = r0, fetch_slot = 0;
        int arrived = r0;       // rows below this one have landed in the ring (always whole pairs)
        auto await = [&](int row, int slot) {
            if (row >= arrived) {
                mbar_wait(&bars[slot >> 1], (phase_bits >> (slot >> 1)) & 1u);
                phase_bits ^= 1u << (slot >> 1);
                arrived += 2;
            }
        };
        auto fetch = [&]() {
            RowTag m = RowTag{0u, 32u};   // past the last row: never plain
            if (fetch_row < r1) {
                await(fetch_row, fetch_slot);
                m = s2_row_meta(ring + fetch_slot * kS2SlotBytes, map_off, lane, lane_relevant);
                map_off = (map_off + map_step) & 7;
                if (++fetch_slot == kS2RingDepth) fetch_slot = 0;
                ++fetch_row;
            }
            return m;
        };
        RowTag m0 = fetch(), m1 = fetch();
        // What the pipeline computes from the rows above r0 (zero state) never reaches an owned row
        // (the dependency cone of rows >= ys at level K ends at row r0 at level 0), so those rows
        // may as well count as rows like the first one: a task whose first rows are steady starts in
        // the branch-free body straight away.
        if (m0.steady()) {
#pragma unroll
            for (int s = 0; s <= K; ++s) info[s] = m0;
            run = K + 1;
        }

        int slot = 0;
        int r = r0;
        // ---- steady pairs: rows r-K-1 .. r+1 carry the same map words (one material, no table
        // lookup, no probe, constant operations on component CC only, or none: CC = -1) -------------
        // (slot even: rows r, r+1 are one ring pair; here fetch_row = r + 2)
        auto steady_pairs = [&](auto cc_tag, auto uni_tag) {
            constexpr int CC = decltype(cc_tag)::value;
            constexpr bool UNI = decltype(uni_tag)::value;   // one material: warp-uniform coefficients
            if (STATS && lane == 0) atomicAdd(a.stats + (UNI ? CC + 1 : 4), 1ull);
            const unsigned my_ids = info[0].ids;
            double gx[C + 1], gy[C], fx[C + 1], fy[C];
            {
                const unsigned left = __shfl_up_sync(0xffffffffu, my_ids, 1);
                const unsigned right = __shfl_down_sync(0xffffffffu, my_ids, 1);
                const unsigned material = info[0].bits & kIdMask;
#pragma unroll
                for (int c = 0; c <= C; ++c) {
                    // cell c-1 (c = 0: last cell of the left-hand lane) and cell c (c = C: first cell
                    // of the right-hand lane)
                    const unsigned below = c ? my_ids >> (16 * (c - 1)) : left >> (16 * (C - 1));
                    const unsigned here = c < C ? my_ids >> (16 * c) : right;
                    gx[c] = tabs[FDS_TAB_GX][UNI ? material : (below & kIdMask)];
                    fx[c] = AXI ? fx_col(UNI ? material : here, c)
                                : tabs[FDS_TAB_FX][UNI ? material : (here & kIdMask)];
                    if (c < C) {
                        gy[c] = tabs[FDS_TAB_GY][UNI ? material : (here & kIdMask)];
                        fy[c] = tabs[FDS_TAB_FY][UNI ? material : (here & kIdMask)];
                    }
                }
            }
            double ca[C], cv[K][C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned k =
                    CC < 0 ? 0u : (my_ids >> (16 * c + class_shift(CC < 0 ? 0 : CC))) & 7u;
                ca[c] = k ? cls_alpha[CC < 0 ? 0 : CC][k] : 1.0;
#pragma unroll
                for (int s = 0; s < K; ++s) cv[s][c] = k ? cls_value[s][CC < 0 ? 0 : CC][k] : -0.0;
            }
            double rv[C + 1];
#pragma unroll
            for (int c = 0; c <= C; ++c) rv[c] = AXI ? r_col(c) : 1.0;
            for (;;) {
                double cur[3][C], pb1[K][C], un1[K][C], vn1[K][C];
                load_row(cur, ring + slot * kS2SlotBytes);
#pragma unroll
                for (int s = 0; s < K; ++s)
                    steady_stage<THERMAL, CC, AXI>(cur, pb[s], un[s], vn[s], pb1[s], un1[s], vn1[s],
                                                   gx, gy, fx, fy, ca, cv[s], rv);
                store_row(cur, r - K, cell_r - K * nx);

                load_row(cur, ring + (slot + 1) * kS2SlotBytes);
                __syncwarp();
                if (lane == 0 && r + kS2RingDepth < r1) issue_pair(r + kS2RingDepth, slot);
#pragma unroll
                for (int s = 0; s < K; ++s)
                    steady_stage<THERMAL, CC, AXI>(cur, pb1[s], un1[s], vn1[s], pb[s], un[s], vn[s],
                                                   gx, gy, fx, fy, ca, cv[s], rv);
                store_row(cur, r + 1 - K, cell_r + nx - K * nx);
                // the window info[0..K] stays what it was: steady rows with these map words
                cell_r += 2 * nx;
                r += 2;
                slot = slot + 2 == kS2RingDepth ? 0 : slot + 2;
                // In here the lookahead is implied: rows r, r+1 have landed and been examined
                // (fetch_row = r + 2); the three cursors are only brought up to date on the way out.
                fetch_row = r;
                fetch_slot = slot;
                arrived = r;
                if (r + 1 < r1) {
                    // next pair: one vote instead of the full metadata
                    mbar_wait(&bars[slot >> 1], (phase_bits >> (slot >> 1)) & 1u);
                    phase_bits ^= 1u << (slot >> 1);
                    arrived = r + 2;
                    const unsigned char *src = ring + slot * kS2SlotBytes + 3 * kS2FieldBytes;
                    const int map_off1 = (map_off + map_step) & 7;
                    const unsigned raw0 =
                        *reinterpret_cast<const unsigned *>(src + 2 * map_off + 4 * lane);
                    const unsigned raw1 = *reinterpret_cast<const unsigned *>(
                        src + kS2SlotBytes + 2 * map_off1 + 4 * lane);
                    if (__all_sync(0xffffffffu,
                                   !lane_relevant || (raw0 == my_ids && raw1 == my_ids))) {
                        map_off = (map_off1 + map_step) & 7;
                        continue;
                    }
                }
                m0 = fetch();
                m1 = fetch();
                break;
            }
        };

        while (r < r1) {
            if (run > K && !(slot & 1) && m0.bits == info[0].bits && m1.bits == m0.bits &&
                (m0.plain() ||
                 __all_sync(0xffffffffu, m0.ids == info[0].ids && m1.ids == info[0].ids))) {
                using std::integral_constant;
                if (m0.bits & 32u)
                    steady_pairs(integral_constant<int, -1>{}, integral_constant<bool, false>{});
                else
                    switch (m0.classed()) {
                        case 0u:
                            steady_pairs(integral_constant<int, -1>{}, integral_constant<bool, true>{});
                            break;
                        case 1u:
                            steady_pairs(integral_constant<int, 0>{}, integral_constant<bool, true>{});
                            break;
                        case 2u:
                            steady_pairs(integral_constant<int, 1>{}, integral_constant<bool, true>{});
                            break;
                        default:
                            steady_pairs(integral_constant<int, 2>{}, integral_constant<bool, true>{});
                            break;
                    }
                continue;
            }

            // ---- general row iteration -----------------------------------------------------------
            if (STATS && lane == 0) atomicAdd(a.stats + 5, 1ull);
            double cur[3][C];
            load_row(cur, ring + slot * kS2SlotBytes);
            __syncwarp();
            // the pair of ring slots is free once its second row has been read
            if (lane == 0 && (slot & 1) && r - 1 + kS2RingDepth < r1)
                issue_pair(r - 1 + kS2RingDepth, slot - 1);

#pragma unroll
            for (int s = K; s > 0; --s) info[s] = info[s - 1];
            info[0] = m0;
            // run = rows in a row, the last one being info[0], that are steady with the same map words
            if (!info[0].steady())
                run = 0;
            else if (run > 0 && info[0].bits == info[1].bits &&
                     (info[0].plain() || __all_sync(0xffffffffu, info[0].ids == info[1].ids)))
                ++run;
            else
                run = 1;

            // The K stages run as a ROLLED loop over one copy of the stage code (this path is rare: code
            // size and register pressure matter more than the moves): stage s works on pb[0], un[0],
            // vn[0] and on info[0], info[1]; after each stage the state and the tags rotate by one
            // place. K rotations put the K state entries back where they belong, the K + 1 tags need
            // one more after the loop.
#pragma unroll 1
            for (int s = 0; s < K; ++s) {
                // stage s: cur = row q = r - s at level s  ->  cur = row q-1 at level s+1
                const RowTag ri = info[0], rp = info[1];
                const int ri_uniform = ri.uniform();
                const int q = r - s;

                // 1. boundaries and probes of p (row q)
                if (ri.classed() & 1u) {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        cur[0][c] = apply_class(cls_alpha, cls_value[s], 0, ri.ids >> (16 * c),
                                                cur[0][c]);
                }
                if (ri.flagged()) {
#pragma unroll
                    for (int c = 0; c < C; ++c) scratch[c] = cur[0][c];
                    s2_slow_cells(a.tables, 0, 1, a.sig_index + s, a.ring_row + s, cell_r - s * nx,
                                  ri.ids, lane_owned && q >= ys && q < ye, scratch);
#pragma unroll
                    for (int c = 0; c < C; ++c) cur[0][c] = scratch[c];
                }

                // 2. new vx, vy of row q (backward differences of p), their boundaries and probes;
                // 3. new p of row q-1 (forward differences of the new vx, vy).
                // `coef` yields the material coefficient of a cell: gx(c) for row q, cell c-1 (c = 0
                // is the left-hand lane's last cell); gyc/fyc(c) row q; gyp/fxp/fyp(c) row q-1
                // (fxp(C) is the right-hand lane's first cell).
                double nu[C], nv[C], np[C];
                auto math = [&](const auto &coef) {
                    const double p_left = shfl_up1(cur[0][C - 1]);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const double pl = c ? cur[0][c - 1] : p_left;
                        const double dx_ = diff2(coef.gx(c), pl, coef.gx(c + 1), cur[0][c]);
                        const double dy_ = diff2(coef.gyp(c), pb[0][c], coef.gyc(c), cur[0][c]);
                        nu[c] = THERMAL ? -dx_
                                : AXI   ? sub(cur[1][c], add(dx_, mul(0.0, cur[1][c])))
                                        : sub(cur[1][c], dx_);
                        nv[c] = THERMAL ? -dy_ : sub(cur[2][c], dy_);
                    }
                    if (ri.classed() & 2u) {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            nu[c] = apply_class(cls_alpha, cls_value[s], 1, ri.ids >> (16 * c), nu[c]);
                    }
                    if (ri.classed() & 4u) {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            nv[c] = apply_class(cls_alpha, cls_value[s], 2, ri.ids >> (16 * c), nv[c]);
                    }
                    if (ri.flagged()) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            scratch[c] = nu[c];
                            scratch[C + c] = nv[c];
                        }
                        s2_slow_cells(a.tables, 1, 2, a.sig_index + s, a.ring_row + s,
                                      cell_r - s * nx, ri.ids, lane_owned && q >= ys && q < ye,
                                      scratch);
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            nu[c] = scratch[c];
                            nv[c] = scratch[C + c];
                        }
                    }
                    const double u_right = shfl_down1(un[0][0]);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        double ul = un[0][c], ur = c < C - 1 ? un[0][c + 1] : u_right;
                        double fxa = coef.fxp(c), fxb = coef.fxp(c + 1);
                        if (AXI) {
                            ul = mul(ul, r_col(c));
                            ur = mul(ur, r_col(c + 1));
                            fxa = fx_col(coef.mfxp(c), c);
                            fxb = fx_col(coef.mfxp(c + 1), c + 1);
                        }
                        const double divx = diff2(fxa, ul, fxb, ur);
                        const double divy = diff2(coef.fyp(c), vn[0][c], coef.fyc(c), nv[c]);
                        np[c] = sub(pb[0][c], add(divx, divy));
                    }
                };
                if (ri_uniform >= 0 && ri_uniform == rp.uniform()) {
                    // all cells of rows q-1 and q share one material: 4 coefficients in registers
                    struct {
                        double g0, g1, f0, f1;
                        unsigned material;
                        __device__ unsigned mfxp(int) const { return material; }
                        __device__ double gx(int) const { return g0; }
                        __device__ double gyc(int) const { return g1; }
                        __device__ double gyp(int) const { return g1; }
                        __device__ double fxp(int) const { return f0; }
                        __device__ double fyp(int) const { return f1; }
                        __device__ double fyc(int) const { return f1; }
                    } coef{tabs[FDS_TAB_GX][ri_uniform], tabs[FDS_TAB_GY][ri_uniform],
                           tabs[FDS_TAB_FX][ri_uniform], tabs[FDS_TAB_FY][ri_uniform],
                           (unsigned)ri_uniform};
                    math(coef);
                } else {
                    struct {
                        const double (*tabs)[kMaxMaterials];
                        unsigned cur_ids, prev_ids, left, right;
                        __device__ int mc(int c) const { return (int)(cur_ids >> (16 * c)) & kIdMask; }
                        __device__ int mp(int c) const { return (int)(prev_ids >> (16 * c)) & kIdMask; }
                        __device__ double gx(int c) const {
                            return tabs[FDS_TAB_GX][c ? mc(c - 1)
                                                      : (int)((left >> (16 * (C - 1))) & kIdMask)];
                        }
                        __device__ double gyc(int c) const { return tabs[FDS_TAB_GY][mc(c)]; }
                        __device__ double gyp(int c) const { return tabs[FDS_TAB_GY][mp(c)]; }
                        __device__ unsigned mfxp(int c) const {
                            return c < C ? (unsigned)mp(c) : (right & kIdMask);
                        }
                        __device__ double fxp(int c) const {
                            return tabs[FDS_TAB_FX][c < C ? mp(c) : (int)(right & kIdMask)];
                        }
                        __device__ double fyp(int c) const { return tabs[FDS_TAB_FY][mp(c)]; }
                        __device__ double fyc(int c) const { return tabs[FDS_TAB_FY][mc(c)]; }
                    } coef{tabs, ri.ids, rp.ids,
                           __shfl_up_sync(0xffffffffu, ri.ids, 1),       // lane-1, row q
                           __shfl_down_sync(0xffffffffu, rp.ids, 1)};    // lane+1, row q-1
                    math(coef);
                }

                // 4. hand row q-1 (level s+1) to the next stage, keep row q for the next iteration,
                //    rotate: stage s+1 finds its state at index 0
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double keep_p = cur[0][c];
                    cur[0][c] = np[c];
                    cur[1][c] = un[0][c];
                    cur[2][c] = vn[0][c];
#pragma unroll
                    for (int j = 0; j + 1 < K; ++j) {
                        pb[j][c] = pb[j + 1][c];
                        un[j][c] = un[j + 1][c];
                        vn[j][c] = vn[j + 1][c];
                    }
                    pb[K - 1][c] = keep_p;
                    un[K - 1][c] = nu[c];
                    vn[K - 1][c] = nv[c];
                }
                {
                    const RowTag first = info[0];
#pragma unroll
                    for (int j = 0; j < K; ++j) info[j] = info[j + 1];
                    info[K] = first;
                }
            }
            {
                const RowTag first = info[0];
#pragma unroll
                for (int j = 0; j < K; ++j) info[j] = info[j + 1];
                info[K] = first;
            }

            store_row(cur, r - K, cell_r - K * nx);   // row r-K at level K
            cell_r += nx;
            ++r;
            if (++slot == kS2RingDepth) slot = 0;
            m0 = m1;
            m1 = fetch();
        }
        task_release(a.sync, a.out, nx, stash[0], (int)xs + kS2StripHalo, ys, ye, stash[1], lane);
    }
}

}  // namespace fds
