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

template <int MODEL, bool LOSSY>
__global__ void __launch_bounds__(256) step2d_kernel(Step2DArgs a, StepTables t, int n_mat1) {
    constexpr bool kAxi = (MODEL == FDS_ACOUSTIC3DAXI || MODEL == FDS_THERMAL3DAXI);
    constexpr bool kThermal = (MODEL == FDS_THERMAL2D || MODEL == FDS_THERMAL3DAXI);
    constexpr bool kVisc = LOSSY && !kThermal;

    const long long x = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const long long nx = a.nx;
    const long long i = (a.row_begin + blockIdx.x) * nx + x;
    const map_t *__restrict__ map = t.map;
    const double *__restrict__ sin_ = a.in[0];
    const double *__restrict__ uin = a.in[1];
    const double *__restrict__ win = a.in[2];

    // ---- all loads first ------------------------------------------------------------------------
    // map bytes / scalar field at the 5-point stencil
    unsigned m0 = map[i], mxm = map[i - 1], mxp = map[i + 1], mym = map[i - nx], myp = map[i + nx];
    double s0 = sin_[i], sxm = sin_[i - 1], sxp = sin_[i + 1], sym = sin_[i - nx],
           syp = sin_[i + nx];
    double u0 = 0, u1 = 0, w0 = 0, w1 = 0;
    if (!kThermal) {
        u0 = uin[i]; u1 = uin[i + 1];
        w0 = win[i]; w1 = win[i + nx];
    }
    // viscous operator: old vector components around cells i, i+1 (x) and i, i+nx (y)
    double ua[2] = {0, 0}, ub[2] = {0, 0}, ul = 0, ur = 0;         // vx at j-nx, j+nx; i-1, i+2
    double wa = 0, wb = 0, wl[2] = {0, 0}, wr[2] = {0, 0};         // vy at i-nx, i+2nx; j-1, j+1
    unsigned mxa[2] = {0, 0}, mxb[2] = {0, 0}, mx2 = 0;            // map at i-nx+k, i+nx+k, i+2
    unsigned myl = 0, my2 = 0;                                     // map at i+nx-1, i+2nx
    if (kVisc) {
        ua[0] = uin[i - nx]; ua[1] = uin[i - nx + 1];
        ub[0] = uin[i + nx]; ub[1] = uin[i + nx + 1];
        ul = uin[i - 1]; ur = uin[i + 2];
        wa = win[i - nx]; wb = win[i + 2 * nx];
        wl[0] = win[i - 1]; wl[1] = win[i + nx - 1];
        wr[0] = win[i + 1]; wr[1] = win[i + nx + 1];
        mxa[0] = mym; mxa[1] = map[i - nx + 1];
        mxb[0] = myp; mxb[1] = map[i + nx + 1];
        mx2 = map[i + 2];
        myl = map[i + nx - 1]; my2 = map[i + 2 * nx];
    }

    auto tab = [&](int which, unsigned m) {
        return __ldg(t.tab + which * kMaxMaterials + (m & kIdMask));
    };
    auto ctab = [&](int which, unsigned m, long long col) {
        return __ldg(t.ctab + ((long long)which * n_mat1 + (m & kIdMask)) * nx + col);
    };

    // boundary operations of one cell and component: the inline constant class, else the table
    auto bound = [&](int comp, unsigned m, long long cell, double v) {
        if (m & (kFlagBound | ((kMaxClasses - 1u) << class_shift(comp)))) {
            v = apply_class(t.cls_alpha, t.cls_value, comp, m, v);
            if (m & kFlagBound)
                v = apply_bounds(t.bound[comp], t.rows, t.signals, t.sig_steps, a.sig_index, cell, v);
        }
        return v;
    };

    // ---- step 1: boundaries and probes of the scalar component ----------------------------------
    double *__restrict__ record = t.ring + a.ring_row * t.n_slots;
    const unsigned any = m0 | mxm | mxp | mym | myp;
    if (any & (kFlagBound | kClassMask)) {
        s0 = bound(0, m0, i, s0);
        sxm = bound(0, mxm, i - 1, sxm);
        sxp = bound(0, mxp, i + 1, sxp);
        sym = bound(0, mym, i - nx, sym);
        syp = bound(0, myp, i + nx, syp);
    }
    if (m0 & kFlagProbe) write_probes(t.probe[0], t.rows, record, i, s0);

    // ---- step 2/3: x component at cells i and i+1 ------------------------------------------------
    double ux[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long j = i + k;
        const unsigned mj = k ? mxp : m0, mjm = k ? m0 : mxm;
        const double sj = k ? sxp : s0, sjm = k ? s0 : sxm;
        // A_vx_p p  |  A_qx_t T : backward difference, offsets [-1, 0]
        const double d = diff2(tab(FDS_TAB_GX, mjm), sjm, tab(FDS_TAB_GX, mj), sj);
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
                    cm1 = tab(FDS_TAB_VM1, mjm);
                    cp1 = tab(FDS_TAB_VP1, mc);
                }
                // V u: five diagonals in offset order [-nx, -1, 0, +1, +nx]
                double vis = acc0(mul(tab(FDS_TAB_VMN, mxa[k]), ua[k]));
                vis = add(vis, mul(cm1, um));
                vis = add(vis, mul(tab(FDS_TAB_V0, mj), old));
                vis = add(vis, mul(cp1, up));
                vis = add(vis, mul(tab(FDS_TAB_VPN, mxb[k]), ub[k]));
                double rhs = sub(d, vis);
                if (kAxi) {
                    // + dt*mu/rho * vx / r**2   (pyfds/acoustics.py:213-215)
                    const double e = mul(tab(FDS_TAB_EB, mj), old) /
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
    if (m0 & kFlagProbe) write_probes(t.probe[1], t.rows, record, i, ux[0]);

    // ---- step 2/3: y component at cells i and i+nx -----------------------------------------------
    double uy[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long j = i + k * nx;
        const unsigned mj = k ? myp : m0, mjm = k ? m0 : mym;
        const double sj = k ? syp : s0, sjm = k ? s0 : sym;
        const double d = diff2(tab(FDS_TAB_GY, mjm), sjm, tab(FDS_TAB_GY, mj), sj);
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
                    cm1 = tab(FDS_TAB_VM1, ml);
                    cp1 = tab(FDS_TAB_VP1, mr);
                }
                double vis = acc0(mul(tab(FDS_TAB_VMN, ma), below));
                vis = add(vis, mul(cm1, wl[k]));
                vis = add(vis, mul(tab(FDS_TAB_V0, mj), old));
                vis = add(vis, mul(cp1, wr[k]));
                vis = add(vis, mul(tab(FDS_TAB_VPN, mb), above));
                v = sub(old, sub(d, vis));
            } else {
                v = sub(old, d);
            }
        }
        uy[k] = bound(2, mj, j, v);
    }
    if (m0 & kFlagProbe) write_probes(t.probe[2], t.rows, record, i, uy[0]);

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
        fx0 = tab(FDS_TAB_FX, m0);
        fx1 = tab(FDS_TAB_FX, mxp);
    }
    const double divx = diff2(fx0, f0, fx1, f1);
    const double divy = diff2(tab(FDS_TAB_FY, m0), uy[0], tab(FDS_TAB_FY, myp), uy[1]);
    a.out[0][i] = sub(s0, add(divx, divy));
    if (!kThermal || a.write_vector) {
        a.out[1][i] = ux[0];
        a.out[2][i] = uy[0];
    }
}

}  // namespace fds
