// One-step 2-D kernels ("reference kernel"): one thread per cell, out of place, one launch per time
// step. They are the general path (any nx, lossy, axisymmetric, thermal) and the on-device yardstick
// for the streaming multi-step kernel in fds_stream2d.cuh, which must reproduce them bit for bit.
//
// One launch executes one complete `sim_step` of the reference:
//   acoustic (pyfds/acoustics.py:111-128, 205-225)      thermal (pyfds/thermal.py:92-107, 160-176)
//     1. boundaries on p, probes on p                     1. boundaries on T, probes on T
//     2. vx -= A_vx_p p - V vx (+E), vy likewise          2. qx = -(A_qx_t T), qy = -(A_qy_t T)
//     3. boundaries/probes on vx, then vy                 3. boundaries/probes on qx, then qy
//     4. p -= A_p_vx vx + A_p_vy vy                       4. T -= A_t_qx qx + A_t_qy qy
// Because step 4 of cell i needs the *new* vector components of cells i+1 and i+nx, those are
// recomputed by the thread of cell i (they are not communicated); every value written is computed
// with the reference's exact operation order, so redundant copies agree bitwise.
//
// All global loads of a thread are issued before the first data-dependent branch (the boundary /
// probe flags): the kernel is latency bound otherwise (ncu, profiles/r1_step2d_v1.md).
#pragma once

#include "fds_common.cuh"

namespace fds {

struct Step2DArgs {
    const double *in[3];   // state before the step, origin at local cell 0
    double *out[3];        // state after the step
    long long nx;
    long long row_begin;   // local rows [row_begin, row_end) are updated
    long long row_end;
    long long sig_index;   // step - sig_first_step
    long long ring_row;    // probe record of this step
    int write_vector;      // thermal: also store qx, qy (only needed when the host reads them)
};

// column of a neighbour cell: rows wrap, the x operators of the reference couple the last cell of a
// row to the first cell of the next (pyfds/fields.py:290-297)
__device__ __forceinline__ long long wrap_col(long long col, long long nx) {
    return col < 0 ? col + nx : (col >= nx ? col - nx : col);
}

// Geometry of the shared-memory tile kernel.
constexpr int kTileW = 128;                       // owned cells per tile row
constexpr int kTileHaloX = 8;                     // extra cells either side (16-byte aligned map rows)
constexpr int kTilePitch = kTileW + 2 * kTileHaloX;
constexpr int kTileThreads = 256;

struct Tile2DArgs {
    int tile_h;          // owned rows per tile
    int rows_below;      // extra rows loaded below / above the owned rows
    int rows_above;
    int tiles_x;
    int n_tiles;
    int *counter;        // dynamic tile distribution (zero before the launch)
};

// One complete `sim_step` for cell i. sp/up/wp/mp point AT the cell in a view of the three state
// components and the map whose rows are `pitch` elements apart (global memory: pitch = nx; the
// shared-memory tile: pitch = kTilePitch); x is the cell's column, i its local flat index (for the
// boundary/probe tables and the output).
template <int MODEL, bool LOSSY, bool FLAGS, bool PERCELL = false>
__device__ __forceinline__ void cell_body(const Step2DArgs &a, const StepTables &t, int n_mat1,
                                            long long i, long long x, long long pitch,
                                            const double *__restrict__ sp,
                                            const double *__restrict__ up,
                                            const double *__restrict__ wp,
                                            const map_t *__restrict__ mp,
                                            const double (*__restrict__ tabs)[kMaxMaterials]) {
    constexpr bool kAxi = (MODEL == FDS_ACOUSTIC3DAXI || MODEL == FDS_THERMAL3DAXI);
    constexpr bool kThermal = (MODEL == FDS_THERMAL2D || MODEL == FDS_THERMAL3DAXI);
    constexpr bool kVisc = LOSSY && !kThermal;
    const long long nx = a.nx;

    // ---- all loads first ------------------------------------------------------------------------
    // map bytes / scalar field at the 5-point stencil
    unsigned m0 = mp[0], mxm = mp[-1], mxp = mp[1], mym = mp[-pitch], myp = mp[pitch];
    double s0 = sp[0], sxm = sp[-1], sxp = sp[1], sym = sp[-pitch], syp = sp[pitch];
    double u0 = 0, u1 = 0, w0 = 0, w1 = 0;
    if (!kThermal) {
        u0 = up[0]; u1 = up[1];
        w0 = wp[0]; w1 = wp[pitch];
    }
    // viscous operator: old vector components around cells i, i+1 (x) and i, i+nx (y)
    double ua[2] = {0, 0}, ub[2] = {0, 0}, ul = 0, ur = 0;         // vx at j-nx, j+nx; i-1, i+2
    double wa = 0, wb = 0, wl[2] = {0, 0}, wr[2] = {0, 0};         // vy at i-nx, i+2nx; j-1, j+1
    unsigned mxa[2] = {0, 0}, mxb[2] = {0, 0}, mx2 = 0;            // map at i-nx+k, i+nx+k, i+2
    unsigned myl = 0, my2 = 0;                                     // map at i+nx-1, i+2nx
    if (kVisc) {
        ua[0] = up[-pitch]; ua[1] = up[1 - pitch];
        ub[0] = up[pitch]; ub[1] = up[pitch + 1];
        ul = up[-1]; ur = up[2];
        wa = wp[-pitch]; wb = wp[2 * pitch];
        wl[0] = wp[-1]; wl[1] = wp[pitch - 1];
        wr[0] = wp[1]; wr[1] = wp[pitch + 1];
        mxa[0] = mym; mxa[1] = mp[1 - pitch];
        mxb[0] = myp; mxb[1] = mp[pitch + 1];
        mx2 = mp[2];
        myl = mp[pitch - 1]; my2 = mp[2 * pitch];
    }

    // coefficient of the cell `off` places from cell i whose map entry is m: by material, or -- where
    // the materials differ from cell to cell (PERCELL, MaterialCoupling on a smooth source) -- from the
    // per-cell arrays; cells outside the grid are void (all coefficients zero) either way
    auto tab = [&](int which, unsigned m, long long off) {
        if (PERCELL) return (m & kIdMask) ? t.cell_tab[(long long)which * t.cell_n + i + off] : 0.0;
        return tabs[which][m & kIdMask];
    };
    auto ctab = [&](int which, unsigned m, long long col) {
        return __ldg(t.ctab + ((long long)which * n_mat1 + (m & kIdMask)) * nx + col);
    };

    // boundary operations of one cell and component: the inline constant class, else the table
    auto bound = [&](int comp, unsigned m, long long cell, double v) {
        if (FLAGS && (m & (kFlagBound | ((kMaxClasses - 1u) << class_shift(comp))))) {
            const unsigned k = (m >> class_shift(comp)) & (kMaxClasses - 1);
            if (k) v = add(mul(t.cls_alpha[comp][k], v), class_value(t, comp, k, a.sig_index));
            if (m & kFlagBound)
                v = apply_bounds(t.bound[comp], t.rows, t.signals, t.sig_steps, a.sig_index, cell, v);
        }
        return v;
    };

    // ---- step 1: boundaries and probes of the scalar component ----------------------------------
    double *__restrict__ record = t.ring + a.ring_row * t.n_slots;
    const unsigned any = m0 | mxm | mxp | mym | myp;
    if (FLAGS && (any & (kFlagBound | kClassMask))) {
        s0 = bound(0, m0, i, s0);
        sxm = bound(0, mxm, i - 1, sxm);
        sxp = bound(0, mxp, i + 1, sxp);
        sym = bound(0, mym, i - nx, sym);
        syp = bound(0, myp, i + nx, syp);
    }
    if (FLAGS && (m0 & kFlagProbe)) write_probes(t.probe[0], t.rows, record, i, s0);

    // ---- step 2/3: x component at cells i and i+1 ------------------------------------------------
    double ux[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long j = i + k;
        const unsigned mj = k ? mxp : m0, mjm = k ? m0 : mxm;
        const double sj = k ? sxp : s0, sjm = k ? s0 : sxm;
        // A_vx_p p  |  A_qx_t T : backward difference, offsets [-1, 0]
        const double d = diff2(tab(FDS_TAB_GX, mjm, k - 1), sjm, tab(FDS_TAB_GX, mj, k), sj);
        double v;
        if (kThermal) {
            v = -d;
        } else {
            const double old = k ? u1 : u0;
            if (kVisc) {
                const long long col = wrap_col(x + k, nx);
                const unsigned mc = k ? mx2 : mxp;           // material of cell j+1
                const double um = k ? u0 : ul, up = k ? ur : u1;
                double cm1, cp1;
                if (kAxi) {
                    cm1 = ctab(FDS_CTAB_VM1, mjm, wrap_col(col - 1, nx));
                    cp1 = ctab(FDS_CTAB_VP1, mc, wrap_col(col + 1, nx));
                } else {
                    cm1 = tab(FDS_TAB_VM1, mjm, k - 1);
                    cp1 = tab(FDS_TAB_VP1, mc, k + 1);
                }
                // V u: five diagonals in offset order [-nx, -1, 0, +1, +nx]
                double vis = acc0(mul(tab(FDS_TAB_VMN, mxa[k], k - nx), ua[k]));
                vis = add(vis, mul(cm1, um));
                vis = add(vis, mul(tab(FDS_TAB_V0, mj, k), old));
                vis = add(vis, mul(cp1, up));
                vis = add(vis, mul(tab(FDS_TAB_VPN, mxb[k], k + nx), ub[k]));
                double rhs = sub(d, vis);
                if (kAxi) {
                    // + dt*mu/rho * vx / r**2   (pyfds/acoustics.py:213-215)
                    const double e = mul(tab(FDS_TAB_EB, mj, k), old) /
                                     __ldg(t.cvec + FDS_CVEC_RR * nx + col);
                    rhs = add(rhs, e);
                }
                v = sub(old, rhs);
            } else if (kAxi) {
                // lossless: V u = +0 and the extra term is (0*vx)/r^2 = a zero with the sign of vx
                v = sub(old, add(d, mul(0.0, old)));
            } else {
                v = sub(old, d);
            }
        }
        ux[k] = bound(1, mj, j, v);
    }
    if (FLAGS && (m0 & kFlagProbe)) write_probes(t.probe[1], t.rows, record, i, ux[0]);

    // ---- step 2/3: y component at cells i and i+nx -----------------------------------------------
    double uy[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long j = i + k * nx;
        const unsigned mj = k ? myp : m0, mjm = k ? m0 : mym;
        const double sj = k ? syp : s0, sjm = k ? s0 : sym;
        const double d = diff2(tab(FDS_TAB_GY, mjm, (k - 1) * nx), sjm, tab(FDS_TAB_GY, mj, k * nx), sj);
        double v;
        if (kThermal) {
            v = -d;
        } else {
            const double old = k ? w1 : w0;
            if (kVisc) {
                // a_vy_vy is the same matrix as a_vx_vx (pyfds/acoustics.py:108,202)
                const unsigned ma = k ? m0 : mym, mb = k ? my2 : myp;
                const unsigned ml = k ? myl : mxm, mr = k ? mxb[1] : mxp;
                const double below = k ? w0 : wa, above = k ? wb : w1;
                double cm1, cp1;
                if (kAxi) {
                    cm1 = ctab(FDS_CTAB_VM1, ml, wrap_col(x - 1, nx));
                    cp1 = ctab(FDS_CTAB_VP1, mr, wrap_col(x + 1, nx));
                } else {
                    cm1 = tab(FDS_TAB_VM1, ml, k * nx - 1);
                    cp1 = tab(FDS_TAB_VP1, mr, k * nx + 1);
                }
                double vis = acc0(mul(tab(FDS_TAB_VMN, ma, (k - 1) * nx), below));
                vis = add(vis, mul(cm1, wl[k]));
                vis = add(vis, mul(tab(FDS_TAB_V0, mj, k * nx), old));
                vis = add(vis, mul(cp1, wr[k]));
                vis = add(vis, mul(tab(FDS_TAB_VPN, mb, (k + 1) * nx), above));
                v = sub(old, sub(d, vis));
            } else {
                v = sub(old, d);
            }
        }
        uy[k] = bound(2, mj, j, v);
    }
    if (FLAGS && (m0 & kFlagProbe)) write_probes(t.probe[2], t.rows, record, i, uy[0]);

    // ---- step 4: scalar update, forward differences with offsets [0, +1] and [0, +nx] -----------
    double fx0, fx1, f0 = ux[0], f1 = ux[1];
    if (kAxi) {
        const long long col1 = wrap_col(x + 1, nx);
        fx0 = ctab(FDS_CTAB_FX, m0, x);
        fx1 = ctab(FDS_CTAB_FX, mxp, col1);
        // the operator is applied to vx * r (pyfds/acoustics.py:224, pyfds/thermal.py:175)
        f0 = mul(f0, __ldg(t.cvec + FDS_CVEC_R * nx + x));
        f1 = mul(f1, __ldg(t.cvec + FDS_CVEC_R * nx + col1));
    } else {
        fx0 = tab(FDS_TAB_FX, m0, 0);
        fx1 = tab(FDS_TAB_FX, mxp, 1);
    }
    const double divx = diff2(fx0, f0, fx1, f1);
    const double divy = diff2(tab(FDS_TAB_FY, m0, 0), uy[0], tab(FDS_TAB_FY, myp, nx), uy[1]);
    a.out[0][i] = sub(s0, add(divx, divy));
    if (!kThermal || a.write_vector) {
        a.out[1][i] = ux[0];
        a.out[2][i] = uy[0];
    }
}

// Cells whose 5-point neighbourhood carries no boundary operation, class or probe (almost all) take a
// body with every flag test compiled out.
template <int MODEL, bool LOSSY, bool PERCELL = false>
__device__ __forceinline__ void cell_update(const Step2DArgs &a, const StepTables &t, int n_mat1,
                                            long long i, long long x, long long pitch,
                                            const double *__restrict__ sp,
                                            const double *__restrict__ up,
                                            const double *__restrict__ wp,
                                            const map_t *__restrict__ mp,
                                            const double (*__restrict__ tabs)[kMaxMaterials]) {
    const unsigned any = mp[0] | mp[-1] | mp[1] | mp[-pitch] | mp[pitch];
    if (any & (kFlagBound | kFlagProbe | kClassMask))
        cell_body<MODEL, LOSSY, true, PERCELL>(a, t, n_mat1, i, x, pitch, sp, up, wp, mp, tabs);
    else
        cell_body<MODEL, LOSSY, false, PERCELL>(a, t, n_mat1, i, x, pitch, sp, up, wp, mp, tabs);
}

// One thread per cell, operands straight from global memory (L1/L2 provide the neighbour reuse).
// PERCELL: coefficients from per-cell arrays (StepTables::cell_tab) instead of the material table.
template <int MODEL, bool LOSSY, bool PERCELL = false>
__global__ void __launch_bounds__(256) step2d_kernel(Step2DArgs a, StepTables t, int n_mat1) {
    __shared__ double tabs[FDS_TAB_COUNT][kMaxMaterials];
    for (int k = threadIdx.x; k < FDS_TAB_COUNT * n_mat1; k += blockDim.x)
        tabs[k / n_mat1][k % n_mat1] = t.tab[(k / n_mat1) * kMaxMaterials + k % n_mat1];
    __syncthreads();
    const long long x = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const long long i = (a.row_begin + blockIdx.x) * a.nx + x;
    cell_update<MODEL, LOSSY, PERCELL>(a, t, n_mat1, i, x, a.nx, a.in[0] + i, a.in[1] + i,
                                       a.in[2] + i, t.map + i, tabs);
}

__device__ __forceinline__ unsigned tile_smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}

// Tile kernel: a CTA stages a (tile_h + halo rows) x 144-cell tile of every input component and of
// the map in shared memory with bulk async copies (TMA, one mbarrier), computes the 128 x tile_h owned
// cells from shared memory with the same `cell_update`, and takes the next tile from an atomic counter.
// Two CTAs per SM: one computes while the other one's copies are in flight, so DRAM latency is hidden
// by bytes in flight in shared memory instead of registers (the one-step kernel above is latency
// bound: profiles/prof_r1_step2d_b.md). Needs nx % 8 == 0 (16-byte aligned map rows).
template <int MODEL, bool LOSSY>
__global__ void __launch_bounds__(kTileThreads, 2)
tile2d_kernel(Step2DArgs a, StepTables t, int n_mat1, Tile2DArgs g) {
    constexpr bool kThermal = (MODEL == FDS_THERMAL2D || MODEL == FDS_THERMAL3DAXI);
    constexpr int kFields = kThermal ? 1 : 3;
    extern __shared__ __align__(16) unsigned char tile_raw[];
    __shared__ unsigned long long bar;
    __shared__ int next_tile[2];
    __shared__ double tabs[FDS_TAB_COUNT][kMaxMaterials];

    const int rows = g.tile_h + g.rows_below + g.rows_above;
    double *fs = reinterpret_cast<double *>(tile_raw);                 // [kFields][rows][pitch]
    map_t *fm = reinterpret_cast<map_t *>(fs + (size_t)kFields * rows * kTilePitch);
    const int tid = threadIdx.x;
    const long long nx = a.nx;

    for (int k = tid; k < FDS_TAB_COUNT * kMaxMaterials; k += kTileThreads)
        (&tabs[0][0])[k] = t.tab[k];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tile_smem_addr(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned parity = 0;
    for (int it = 0;; ++it) {
        if (tid == 0) next_tile[it & 1] = atomicAdd(g.counter, 1);
        __syncthreads();                 // also: every thread is done with the previous tile
        const int tile = next_tile[it & 1];
        if (tile >= g.n_tiles) break;
        const long long x0 = (long long)(tile % g.tiles_x) * kTileW;
        const long long y0 = a.row_begin + (long long)(tile / g.tiles_x) * g.tile_h;

        if (tid < 32) {
            // generic-proxy reads of the previous tile are ordered before the async-proxy writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tid == 0) {
                const unsigned bytes = (unsigned)rows * (kFields * kTilePitch * 8 + kTilePitch * 2);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                                 tile_smem_addr(&bar)),
                             "r"(bytes)
                             : "memory");
            }
            __syncwarp();
            for (int r = tid; r < rows; r += 32) {
                const long long base = (y0 - g.rows_below + r) * nx + x0 - kTileHaloX;
#pragma unroll
                for (int f = 0; f < kFields; ++f)
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
                        "[%1], %2, [%3];" ::"r"(tile_smem_addr(fs + ((size_t)f * rows + r) * kTilePitch)),
                        "l"(a.in[f] + base), "r"(kTilePitch * 8), "r"(tile_smem_addr(&bar))
                        : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], "
                    "%2, [%3];" ::"r"(tile_smem_addr(fm + (size_t)r * kTilePitch)),
                    "l"(t.map + base), "r"(kTilePitch * 2), "r"(tile_smem_addr(&bar))
                    : "memory");
            }
        }
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "TWAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra TDONE_%=;\n"
            "bra TWAIT_%=;\n"
            "TDONE_%=:\n"
            "}\n" ::"r"(tile_smem_addr(&bar)),
            "r"(parity)
            : "memory");
        parity ^= 1u;

        const int cells = kTileW * g.tile_h;
        for (int k = tid; k < cells; k += kTileThreads) {
            const int lx = k % kTileW, ly = k / kTileW;
            const long long x = x0 + lx, row = y0 + ly;
            if (x >= nx || row >= a.row_end) continue;
            const size_t li = (size_t)(ly + g.rows_below) * kTilePitch + lx + kTileHaloX;
            const double *sp = fs + li;
            const double *up = kThermal ? sp : fs + (size_t)rows * kTilePitch + li;
            const double *wp = kThermal ? sp : fs + (size_t)2 * rows * kTilePitch + li;
            cell_update<MODEL, LOSSY>(a, t, n_mat1, row * nx + x, x, kTilePitch, sp, up, wp, fm + li,
                                      tabs);
        }
    }
}

}  // namespace fds
