// Register-resident 1-D kernel (Acoustic1D pyfds/acoustics.py:40-52, Thermal1D pyfds/thermal.py:40-51).
//
// A 1-D step is a chain of three short dependent updates; with the line in shared memory that chain is
// paid once per block barrier (fds_step1d.cuh: ~1.5 us per step whatever the tile size). Here ONE WARP
// owns a tile of 256 cells for a whole launch: every lane keeps 8 consecutive cells of both components
// and their material coefficients in registers, x-neighbours across lanes travel by warp shuffle, and
// there is no barrier of any kind in the time loop. Warps do not talk to each other (a CTA is just four
// of them, one per SM sub-partition): a tile carries
// `halo` extra cells either side, the region of valid values shrinks by two cells per side and step
// (reach of the viscous operator), so a launch advances halo / 2 steps and the warp stores the
// 256 - 2 * halo cells it owns. A line of at most 256 cells has no neighbours and takes any number of
// steps in one launch.
//
// Arithmetic: the same __dmul_rn/__dadd_rn sequence per cell as step1d_kernel (and the reference's DIA
// mat-vec order). Boundary operations and probes come from the same tables, but a launch is only as
// fast as its slowest warp and the table lookup (an integer division and seven dependent loads that
// mostly come from L2, ~250 cycles each) would cost the warp that holds a source more per step than the
// step itself: the lookups are therefore resolved ONCE per launch into a shared-memory entry per
// flagged cell and component (first operation with its coefficients, operation and probe ranges); in
// the time loop a flagged cell costs one shared-memory read, and one global read if its value is a
// signal. Lanes without a flagged cell skip all of it with one predicate per phase.
#pragma once

#include "fds_common.cuh"
#include "fds_step1d.cuh"

namespace fds {

constexpr int kLineCells = 8;                    // cells per lane
constexpr int kLineWidth = 32 * kLineCells;      // cells per warp tile

// Boundary operations and probes of one cell and component, resolved from the tables.
struct LineEntry {
    double alpha, value;          // first operation: v = alpha * v + (signal sample or value)
    double cls_alpha, cls_value;  // inline class operation, applied first (apply_class)
    long long sigoff;             // signal row offset of the first operation, -1 = scalar value
    int o0, n_ops;                // all operations: o0 .. o0 + n_ops - 1
    int p0, p1;                   // probe entries p0 .. p1 - 1
    int cache;                    // row of the shared-memory signal window, -1 = read from global
    int has_cls;
};

constexpr int kLineSigCache = 16;    // signals whose samples of the next steps are staged on chip
constexpr int kLineSigWindow = 56;   // steps per staging window (>= steps per launch of a tiled line)

// Everything the flagged-cell path needs, in shared memory so that it can live out of line.
struct LineShared {
    LineEntry entries[2][kLineCells][32];
    double sigcache[kLineSigCache][kLineSigWindow];
    long long cache_sigoff[kLineSigCache];   // signal row offset behind each staged row
    const int *bsignal[2];
    const double *balpha[2], *bvalue[2];
    const int *pslots[2];
    const double *signals;
    long long sig_steps;
};

// Entries of one flagged cell (both components) from the host-resolved table.
__device__ __noinline__ void line_resolve(LineShared *sh, int c, int lane,
                                          const int *__restrict__ line_index,
                                          const LineResolved *__restrict__ line_entries,
                                          long long cell, double cls_alpha0, double cls_value0,
                                          int has_cls0, double cls_alpha1, double cls_value1,
                                          int has_cls1) {
    const int number = __ldg(line_index + cell);
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
        LineEntry e{1.0, 0.0, comp ? cls_alpha1 : cls_alpha0, comp ? cls_value1 : cls_value0,
                    -1, 0, 0, 0, 0, -1, comp ? has_cls1 : has_cls0};
        if (number >= 0) {
            const LineResolved r = line_entries[2 * number + comp];
            e.alpha = r.alpha;
            e.value = r.value;
            e.sigoff = r.signal >= 0 ? (long long)r.signal * sh->sig_steps : -1;
            e.o0 = r.o0;
            e.n_ops = r.n_ops;
            e.p0 = r.p0;
            e.p1 = r.p1;
        }
        sh->entries[comp][c][lane] = e;
    }
}

// Stages the samples of the next `count` steps of all cached signals; the whole warp copies.
__device__ __noinline__ void line_refill(LineShared *sh, int lane, int rows, long long sig,
                                         int count) {
    for (int r = 0; r < rows; ++r) {
        const double *__restrict__ src = sh->signals + sh->cache_sigoff[r] + sig;
        for (int j = lane; j < count; j += 32) sh->sigcache[r][j] = __ldg(src + j);
    }
    __syncwarp();
}

// Boundary operations in list order (pyfds/regions.py:136-145), then the probes, of one flagged cell.
__device__ __noinline__ double line_special(const LineShared *sh, int comp, int c, int lane,
                                            long long sig, int window_step, double *record,
                                            bool owned, double v) {
    const LineEntry &e = sh->entries[comp][c][lane];
    if (e.has_cls) v = add(mul(e.cls_alpha, v), e.cls_value);
    if (e.n_ops > 0) {
        const double val = e.cache >= 0    ? sh->sigcache[e.cache][window_step]
                           : e.sigoff >= 0 ? __ldg(sh->signals + e.sigoff + sig)
                                           : e.value;
        v = add(mul(e.alpha, v), val);
        for (int o = e.o0 + 1; o < e.o0 + e.n_ops; ++o) {
            const int sidx = __ldg(sh->bsignal[comp] + o);
            const double w = sidx >= 0 ? __ldg(sh->signals + (long long)sidx * sh->sig_steps + sig)
                                       : __ldg(sh->bvalue[comp] + o);
            v = add(mul(__ldg(sh->balpha[comp] + o), v), w);
        }
    }
    if (owned)
        for (int k = e.p0; k < e.p1; ++k) record[__ldg(sh->pslots[comp] + k)] = v;
    return v;
}

constexpr int kLineWarps = 4;        // warps (= tiles) per CTA: one per SM sub-partition

template <bool THERMAL, bool LOSSY>
__device__ __forceinline__ void line1d_body(const Step1DArgs &a, const StepTables &t, int n_tiles) {
    constexpr int C = kLineCells;
    extern __shared__ __align__(16) unsigned char line_smem[];
    // Launched with programmatic stream serialisation (the one-step launches of a coupled group), the
    // grid is scheduled while its predecessor still runs and waits here for it to complete; both
    // instructions are no-ops in an ordinary launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kLineWarps + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;         // warps are independent: no block-level barrier anywhere
    LineShared &sh = reinterpret_cast<LineShared *>(line_smem)[threadIdx.x >> 5];
    const long long origin = (long long)tile * a.tile - a.halo;         // global cell of local 0
    const long long g0 = origin + lane * C;                             // this lane's first cell

    // state, map entries and coefficients of this lane's cells
    double s[C], u[C];
    unsigned id[C];
    double gx[C], fx[C], v0[C], vm1[C], vp1[C];
    unsigned special = 0, real = 0;      // bit c: cell c has boundary operations / probes; is inside
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const long long g = g0 + c;
        s[c] = a.in[0][g];
        u[c] = THERMAL ? 0.0 : a.in[1][g];
        id[c] = t.map[g];
        const int m = id[c] & kIdMask;
        // coefficients by material, or by cell where the materials vary from cell to cell (cells
        // outside the line are void: all coefficients zero). The per-cell arrays are rewritten between
        // launches by the material couplings: plain loads, not the read-only path.
        auto coefficient = [&](int table) {
            if (t.cell_tab) return m ? t.cell_tab[(long long)table * t.cell_n + g] : 0.0;
            return __ldg(t.tab + table * kMaxMaterials + m);
        };
        gx[c] = coefficient(FDS_TAB_GX);
        fx[c] = coefficient(FDS_TAB_FX);
        if (LOSSY) {
            v0[c] = coefficient(FDS_TAB_V0);
            vm1[c] = coefficient(FDS_TAB_VM1);
            vp1[c] = coefficient(FDS_TAB_VP1);
        }
        if (id[c] & (kFlagBound | kFlagProbe | kClassMask)) special |= 1u << c;
        if (m) real |= 1u << c;
    }
    // coefficients of the cells next to this lane's segment (they sit on the neighbour's column)
    const double gx_left = __shfl_up_sync(0xffffffffu, gx[C - 1], 1);
    const double fx_right = __shfl_down_sync(0xffffffffu, fx[0], 1);
    double vm1_left = 0.0, vp1_right = 0.0;
    if (LOSSY) {
        vm1_left = __shfl_up_sync(0xffffffffu, vm1[C - 1], 1);
        vp1_right = __shfl_down_sync(0xffffffffu, vp1[0], 1);
    }
    const bool owned = lane * C >= a.halo && lane * C < a.halo + a.tile;
    const bool edge = real != (1u << C) - 1u;    // some of this lane's cells lie outside the line

    // resolve the flagged cells once: table positions, first operation, class operation
    if (lane == 0) {
        for (int comp = 0; comp < 2; ++comp) {
            sh.bsignal[comp] = t.bound[comp].signal;
            sh.balpha[comp] = t.bound[comp].alpha;
            sh.bvalue[comp] = t.bound[comp].value;
            sh.pslots[comp] = t.probe[comp].slots;
        }
        sh.signals = t.signals;
        sh.sig_steps = t.sig_steps;
    }
    __syncwarp();
    int n_cached = 0;
    if (special) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (!(special >> c & 1u)) continue;
            const unsigned k0 = (id[c] >> class_shift(0)) & (kMaxClasses - 1);
            const unsigned k1 = (id[c] >> class_shift(1)) & (kMaxClasses - 1);
            line_resolve(&sh, c, lane, t.line_index, t.line_entries, g0 + c, t.cls_alpha[0][k0],
                         t.cls_value[0][k0], k0 != 0, t.cls_alpha[1][k1], t.cls_value[1][k1], k1 != 0);
#pragma unroll
            for (int comp = 0; comp < 2; ++comp)
                if (sh.entries[comp][c][lane].sigoff >= 0) ++n_cached;
        }
    }
    // rows of the signal window: exclusive prefix sum of the lanes' demands
    int first_row = n_cached;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, first_row, d);
        if (lane >= d) first_row += up;
    }
    first_row -= n_cached;
    if (special) {
        for (int comp = 0; comp < 2; ++comp)
            for (int c = 0; c < C; ++c) {
                if (!(special >> c & 1u)) continue;
                LineEntry &e = sh.entries[comp][c][lane];
                if (e.sigoff >= 0 && first_row < kLineSigCache) {
                    sh.cache_sigoff[first_row] = e.sigoff;
                    e.cache = first_row++;
                }
            }
    }
    // cells that have something to do per component (a wall on the velocity must not cost the pressure
    // phase a call)
    unsigned todo[2] = {0u, 0u};
    if (special) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (!(special >> c & 1u)) continue;
#pragma unroll
            for (int comp = 0; comp < 2; ++comp) {
                const LineEntry &e = sh.entries[comp][c][lane];
                if (e.has_cls || e.n_ops > 0 || e.p1 > e.p0) todo[comp] |= 1u << c;
            }
        }
    }
    // A flagged cell goes through the out-of-line path by a jump on its index (the state lives in
    // registers, so the index must be static at the call site): one indirect branch per flagged cell
    // instead of a test per cell, which matters because a launch waits for its slowest warp.
#define FDS_LINE_CASE(K)                                                                        \
    case K:                                                                                     \
        x[K] = line_special(&sh, comp, K, lane, sig, window_step, record, owned, x[K]);         \
        break;
    auto run_special = [&](double(&x)[C], unsigned mask, int comp, long long sig, int window_step,
                           double *record) {
#pragma unroll 1
        while (mask) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1;
            switch (c) {
                FDS_LINE_CASE(0) FDS_LINE_CASE(1) FDS_LINE_CASE(2) FDS_LINE_CASE(3)
                FDS_LINE_CASE(4) FDS_LINE_CASE(5) FDS_LINE_CASE(6) FDS_LINE_CASE(7)
            }
        }
    };
#undef FDS_LINE_CASE
    static_assert(kLineCells == 8, "run_special lists the cells explicitly");
    // rows in use by the whole warp (the last lane's inclusive prefix sum)
    const int cache_rows = min(kLineSigCache, __shfl_sync(0xffffffffu, first_row, 31));
    __syncwarp();

    int window_step = kLineSigWindow;    // position inside the staged signal window; full = refill
    for (int q = 0; q < a.n_steps; ++q) {
        const long long sig = a.sig_index + q;
        double *__restrict__ record = t.ring + (a.ring_row + q) * t.n_slots;
        if (window_step == kLineSigWindow) {
            window_step = 0;
            if (cache_rows) line_refill(&sh, lane, cache_rows, sig, min(kLineSigWindow, a.n_steps - q));
        }

        // 1. boundaries and probes of the scalar component
        if (todo[0]) run_special(s, todo[0], 0, sig, window_step, record);

        // 2. vector component: backward difference of the scalar (+ viscous second difference)
        const double s_left = __shfl_up_sync(0xffffffffu, s[C - 1], 1);
        double u_left = 0.0, u_right = 0.0;
        if (LOSSY) {
            u_left = __shfl_up_sync(0xffffffffu, u[C - 1], 1);
            u_right = __shfl_down_sync(0xffffffffu, u[0], 1);
        }
        double unew[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double d = diff2(c ? gx[c - 1] : gx_left, c ? s[c - 1] : s_left, gx[c], s[c]);
            if (THERMAL) {
                unew[c] = -d;
            } else if (LOSSY) {
                double vis = acc0(mul(c ? vm1[c - 1] : vm1_left, c ? u[c - 1] : u_left));
                vis = add(vis, mul(v0[c], u[c]));
                vis = add(vis, mul(c < C - 1 ? vp1[c + 1] : vp1_right, c < C - 1 ? u[c + 1] : u_right));
                unew[c] = sub(u[c], sub(d, vis));
            } else {
                unew[c] = sub(u[c], d);
            }
        }
        if (edge) {                                       // cells outside the line stay as they are
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (!(real >> c & 1u)) unew[c] = u[c];
        }

        // 3. boundaries and probes of the vector component
        if (todo[1]) run_special(unew, todo[1], 1, sig, window_step, record);
#pragma unroll
        for (int c = 0; c < C; ++c) u[c] = unew[c];

        // 4. scalar component: forward difference of the vector component
        const double un_right = __shfl_down_sync(0xffffffffu, u[0], 1);
        double snew[C];
#pragma unroll
        for (int c = 0; c < C; ++c)
            snew[c] = sub(s[c], diff2(fx[c], u[c], c < C - 1 ? fx[c + 1] : fx_right,
                                      c < C - 1 ? u[c + 1] : un_right));
        if (edge) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (real >> c & 1u) s[c] = snew[c];
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) s[c] = snew[c];
        }
        ++window_step;
    }

    if (owned) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const long long g = g0 + c;
            if (g < a.n) {
                a.out[0][g] = s[c];
                a.out[1][g] = u[c];
            }
        }
    }
}

template <bool THERMAL, bool LOSSY>
__global__ void __launch_bounds__(32 * kLineWarps, 1) line1d_kernel(Step1DArgs a, StepTables t,
                                                                    int n_tiles) {
    line1d_body<THERMAL, LOSSY>(a, t, n_tiles);
}

// Two member fields of a SynchronizedFields group in ONE launch: within a common step the members do
// not depend on each other (only the interactions that follow couple them, pyfds/coupling.py:81-87),
// so blockIdx.y picks the field and both lines advance side by side. KIND: 0 acoustic lossless,
// 1 acoustic lossy, 2 thermal.
struct LinePairArgs {
    Step1DArgs a[2];
    int n_tiles[2];
};

template <int KIND>
__device__ __forceinline__ void line1d_member(const Step1DArgs &a, const StepTables &t, int n_tiles) {
    if (KIND == 2) line1d_body<true, false>(a, t, n_tiles);
    else if (KIND == 1) line1d_body<false, true>(a, t, n_tiles);
    else line1d_body<false, false>(a, t, n_tiles);
}

template <int KIND0, int KIND1>
__global__ void __launch_bounds__(32 * kLineWarps, 1)
line1d_pair_kernel(LinePairArgs p, const StepTables *__restrict__ tables) {
    if (blockIdx.y == 0) line1d_member<KIND0>(p.a[0], tables[0], p.n_tiles[0]);
    else line1d_member<KIND1>(p.a[1], tables[1], p.n_tiles[1]);
}

}  // namespace fds
