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

template <int MODEL, bool LOSSY>
struct Cell2D {
    static constexpr bool kAxi = (MODEL == FDS_ACOUSTIC3DAXI || MODEL == FDS_THERMAL3DAXI);
    static constexpr bool kThermal = (MODEL == FDS_THERMAL2D || MODEL == FDS_THERMAL3DAXI);

    const Step2DArgs &a;
    const StepTables &t;
    long long nx;

    __device__ __forceinline__ Cell2D(const Step2DArgs &a_, const StepTables &t_)
        : a(a_), t(t_), nx(a_.nx) {}

    __device__ __forceinline__ double tab(int which, uint8_t m) const {
        return __ldg(t.tab + which * kMaxMaterials + (m & kIdMask));
    }
    // (material, column) table of the axisymmetric models; `n_cols` stride is nx
    __device__ __forceinline__ double ctab(int which, int n_mat1, uint8_t m, long long col) const {
        return __ldg(t.ctab + ((long long)which * n_mat1 + (m & kIdMask)) * nx + col);
    }

    // scalar component after its boundaries (step 1)
    __device__ __forceinline__ double scalar_at(long long j) const {
        double v = a.in[0][j];
        if (t.map[j] & kFlagBound)
            v = apply_bounds(t.bound[0], t.signals, t.sig_steps, a.sig_index, j, v);
        return v;
    }
};

// column of cell j given the column of cell i and j - i in {-1, 0, +1} (rows wrap: the x operators
// of the reference couple the last cell of a row to the first of the next, pyfds/fields.py:290-297)
__device__ __forceinline__ long long wrap_col(long long col, long long nx) {
    return col < 0 ? col + nx : (col >= nx ? col - nx : col);
}

template <int MODEL, bool LOSSY>
__global__ void __launch_bounds__(256) step2d_kernel(Step2DArgs a, StepTables t, int n_mat1) {
    using C = Cell2D<MODEL, LOSSY>;
    const long long x = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const long long row = a.row_begin + blockIdx.x;
    const long long nx = a.nx;
    const long long i = row * nx + x;
    const C c(a, t);
    const uint8_t *__restrict__ map = t.map;

    const uint8_t m0 = map[i], mxm = map[i - 1], mxp = map[i + 1], mym = map[i - nx],
                  myp = map[i + nx];
    // scalar field after boundaries at the five points of the stencil
    const double s0 = c.scalar_at(i), sxm = c.scalar_at(i - 1), sxp = c.scalar_at(i + 1),
                 sym = c.scalar_at(i - nx), syp = c.scalar_at(i + nx);
    double *__restrict__ record = t.ring + a.ring_row * t.n_slots;
    if (m0 & kFlagProbe) write_probes(t.probe[0], record, i, s0);

    const double *__restrict__ uin = a.in[1];
    const double *__restrict__ win = a.in[2];

    // ---- x component at cells i and i+1 --------------------------------------------------------
    double ux[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long j = i + k;
        const uint8_t mj = k ? mxp : m0, mjm = k ? m0 : mxm;
        const double sj = k ? sxp : s0, sjm = k ? s0 : sxm;
        // A_vx_p p  |  A_qx_t T : backward difference, offsets [-1, 0]
        const double d = diff2(c.tab(FDS_TAB_GX, mjm), sjm, c.tab(FDS_TAB_GX, mj), sj);
        double v;
        if (C::kThermal) {
            v = -d;
        } else {
            const double old = uin[j];
            if (LOSSY) {
                const long long col = wrap_col(x + k, nx);
                const uint8_t ma = map[j - nx], mb = map[j + nx], mc = map[j + 1];
                double cm1, cp1;
                if (C::kAxi) {
                    cm1 = c.ctab(FDS_CTAB_VM1, n_mat1, mjm, wrap_col(col - 1, nx));
                    cp1 = c.ctab(FDS_CTAB_VP1, n_mat1, mc, wrap_col(col + 1, nx));
                } else {
                    cm1 = c.tab(FDS_TAB_VM1, mjm);
                    cp1 = c.tab(FDS_TAB_VP1, mc);
                }
                // V u: five diagonals in offset order [-nx, -1, 0, +1, +nx]
                double vis = acc0(mul(c.tab(FDS_TAB_VMN, ma), uin[j - nx]));
                vis = add(vis, mul(cm1, uin[j - 1]));
                vis = add(vis, mul(c.tab(FDS_TAB_V0, mj), old));
                vis = add(vis, mul(cp1, uin[j + 1]));
                vis = add(vis, mul(c.tab(FDS_TAB_VPN, mb), uin[j + nx]));
                double rhs = sub(d, vis);
                if (C::kAxi) {
                    // + dt*mu/rho * vx / r**2   (pyfds/acoustics.py:213-215)
                    const double e = mul(c.tab(FDS_TAB_EB, mj), old) /
                                     __ldg(t.cvec + FDS_CVEC_RR * nx + col);
                    rhs = add(rhs, e);
                }
                v = sub(old, rhs);
            } else if (C::kAxi) {
                // lossless: V u = +0 and the extra term is (0*vx)/r^2 = a zero with the sign of vx
                v = sub(old, add(d, mul(0.0, old)));
            } else {
                v = sub(old, d);
            }
        }
        const uint8_t fj = k ? mxp : m0;
        if (fj & kFlagBound)
            v = apply_bounds(t.bound[1], t.signals, t.sig_steps, a.sig_index, j, v);
        ux[k] = v;
    }
    if (m0 & kFlagProbe) write_probes(t.probe[1], record, i, ux[0]);

    // ---- y component at cells i and i+nx -------------------------------------------------------
    double uy[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long j = i + k * nx;
        const uint8_t mj = k ? myp : m0, mjm = k ? m0 : mym;
        const double sj = k ? syp : s0, sjm = k ? s0 : sym;
        const double d = diff2(c.tab(FDS_TAB_GY, mjm), sjm, c.tab(FDS_TAB_GY, mj), sj);
        double v;
        if (C::kThermal) {
            v = -d;
        } else {
            const double old = win[j];
            if (LOSSY) {
                // a_vy_vy is the same matrix as a_vx_vx (pyfds/acoustics.py:108,202)
                const uint8_t ma = map[j - nx], mb = map[j + nx], ml = map[j - 1], mr = map[j + 1];
                double cm1, cp1;
                if (C::kAxi) {
                    cm1 = c.ctab(FDS_CTAB_VM1, n_mat1, ml, wrap_col(x - 1, nx));
                    cp1 = c.ctab(FDS_CTAB_VP1, n_mat1, mr, wrap_col(x + 1, nx));
                } else {
                    cm1 = c.tab(FDS_TAB_VM1, ml);
                    cp1 = c.tab(FDS_TAB_VP1, mr);
                }
                double vis = acc0(mul(c.tab(FDS_TAB_VMN, ma), win[j - nx]));
                vis = add(vis, mul(cm1, win[j - 1]));
                vis = add(vis, mul(c.tab(FDS_TAB_V0, mj), old));
                vis = add(vis, mul(cp1, win[j + 1]));
                vis = add(vis, mul(c.tab(FDS_TAB_VPN, mb), win[j + nx]));
                v = sub(old, sub(d, vis));
            } else {
                v = sub(old, d);
            }
        }
        if (mj & kFlagBound)
            v = apply_bounds(t.bound[2], t.signals, t.sig_steps, a.sig_index, j, v);
        uy[k] = v;
    }
    if (m0 & kFlagProbe) write_probes(t.probe[2], record, i, uy[0]);

    // ---- scalar update: forward differences, offsets [0, +1] and [0, +nx] ----------------------
    double fx0, fx1, w0 = ux[0], w1 = ux[1];
    if (C::kAxi) {
        const long long col1 = wrap_col(x + 1, nx);
        fx0 = c.ctab(FDS_CTAB_FX, n_mat1, m0, x);
        fx1 = c.ctab(FDS_CTAB_FX, n_mat1, mxp, col1);
        // the operator is applied to vx * r (pyfds/acoustics.py:224, pyfds/thermal.py:175)
        w0 = mul(w0, __ldg(t.cvec + FDS_CVEC_R * nx + x));
        w1 = mul(w1, __ldg(t.cvec + FDS_CVEC_R * nx + col1));
    } else {
        fx0 = c.tab(FDS_TAB_FX, m0);
        fx1 = c.tab(FDS_TAB_FX, mxp);
    }
    const double divx = diff2(fx0, w0, fx1, w1);
    const double divy = diff2(c.tab(FDS_TAB_FY, m0), uy[0], c.tab(FDS_TAB_FY, myp), uy[1]);
    a.out[0][i] = sub(s0, add(divx, divy));
    if (!C::kThermal || a.write_vector) {
        a.out[1][i] = ux[0];
        a.out[2][i] = uy[0];
    }
}

}  // namespace fds
