// Streaming kernel for the viscous (lossy) Acoustic2D leapfrog and for Acoustic3DAxi (lossless and
// lossy): pyfds/acoustics.py:111-128 and 205-225. Same machinery as fds_stream2d.cuh (one warp = one
// 120-cell strip x a chunk of rows, rows through a TMA ring, K-stage time pipeline in registers, lanes
// hold 4 consecutive cells, x-neighbours by shuffle), but the viscous 5-point operator on the OLD
// velocities needs a three-row window, so a stage lags TWO rows behind its input:
//
//   row q arrives at level s  ->  new vx, vy of row q-1 (needs old vx, vy of rows q-2, q-1, q and p of
//   rows q-2, q-1)            ->  new p of row q-2 (needs new vx of row q-2, new vy of rows q-2, q-1)
//
// and the dependency cone grows 1 cell to the left and 2 to the right per step, which the 4-cell strip
// halo covers for K <= 2 (only K = 1 is instantiated: a second stage spills registers and is slower). Per stage 8 row fragments stay in registers (p after boundaries, old vx, vy
// of two rows; new vx, vy of one row). Axisymmetric coefficients depend on the column: for the
// warp-uniform material path the per-column values of this lane's cells (and its two neighbours) are
// kept in registers and reloaded only when the material changes; K = 1 there (register budget).
//
// Every value is produced by the same __dmul_rn/__dadd_rn sequence as cell_body in fds_step2d.cuh.
#pragma once

#include "fds_common.cuh"
#include "fds_stream2d.cuh"

namespace fds {

struct StreamVArgs {
    Stream2DArgs base;
    const double *ctab;   // [FDS_CTAB_COUNT][n_mat1][nx]   (axisymmetric)
    const double *cvec;   // [FDS_CVEC_COUNT][nx]
    int n_mat1;
};

template <int K, bool AXI, bool VISC>
__global__ void __launch_bounds__(kStreamWarps * 32, 2) streamv_kernel(StreamVArgs av) {
    const Stream2DArgs &a = av.base;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double tabs[FDS_TAB_COUNT][kMaxMaterials];
    __shared__ double cls_alpha[3][kMaxClasses], cls_value[3][kMaxClasses];

    for (int k = threadIdx.x; k < FDS_TAB_COUNT * kMaxMaterials; k += blockDim.x)
        (&tabs[0][0])[k] = a.tab[k];
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
    const int swap = (lane >> 2) & 1;
    const int off_a = lane * 32 + swap * 16, off_b = lane * 32 + (swap ^ 1) * 16;
    unsigned phase_bits = 0;
    constexpr int kLag = 2 * K;     // rows between the input row and the row stored

    if (lane == 0) {
        for (int d = 0; d < kRingDepth; ++d) mbar_init(&bars[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(a.task_counter, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const int4 tk = __ldg(a.tasks + task);
        const long long ys = tk.y, ye = tk.z;
        const long long xs = (long long)tk.x * kStripStride - kStripHalo;
        const long long r0 = ys - kLag, r1 = ye + kLag;

        auto issue = [&](long long r, int slot) {
            const long long base = r * nx + xs;
            const long long base8 = base & ~7LL;
            unsigned char *dst = ring + slot * kSlotBytes;
            mbar_expect_tx(&bars[slot], 3 * kStripCells * 8 + kMapWindowBytes);
            bulk_load(dst, a.in[0] + base, kStripCells * 8, &bars[slot]);
            bulk_load(dst + kStripCells * 8, a.in[1] + base, kStripCells * 8, &bars[slot]);
            bulk_load(dst + 2 * kStripCells * 8, a.in[2] + base, kStripCells * 8, &bars[slot]);
            bulk_load(dst + 3 * kStripCells * 8, a.map + base8, kMapWindowBytes, &bars[slot]);
        };
        if (lane == 0)
            for (int d = 0; d < kRingDepth && r0 + d < r1; ++d) issue(r0 + d, d);

        // per stage: p after boundaries, old vx, old vy of rows q-1 [0] and q-2 [1]; new vx, vy of
        // row q-2
        double pb[K][2][4], uo[K][2][4], vo[K][2][4], un[K][4], vn[K][4];
        RowInfo info[2 * K + 1];
#pragma unroll
        for (int s = 0; s < K; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                pb[s][0][c] = pb[s][1][c] = uo[s][0][c] = uo[s][1][c] = 0.0;
                vo[s][0][c] = vo[s][1][c] = un[s][c] = vn[s][c] = 0.0;
            }
#pragma unroll
        for (int s = 0; s <= 2 * K; ++s) info[s] = RowInfo{0ull, 0, false, 0u};

        const long long x0 = xs + 4 * lane;
        const bool lane_owned = lane >= 1 && lane <= 30 && x0 < nx;
        long long cell_r = r0 * nx + x0;
        const int map_step = (int)(nx & 7LL);
        int map_off = (int)((r0 * nx + xs) & 7LL);

        // axisymmetric: per-column values of the uniform material for columns x0-1 .. x0+4
        double cx_fx[5], cx_vm1[4], cx_vp1[4], cx_r[5], cx_rr[4];
        long long cols[6];
        int cx_material = -1;
        if (AXI) {
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                long long col = (x0 - 1 + c) % nx;
                cols[c] = col < 0 ? col + nx : col;
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) cx_r[c] = __ldg(av.cvec + FDS_CVEC_R * nx + cols[c + 1]);
#pragma unroll
            for (int c = 0; c < 4; ++c) cx_rr[c] = __ldg(av.cvec + FDS_CVEC_RR * nx + cols[c + 1]);
        }
        auto ctab = [&](int which, int m, long long col) {
            return __ldg(av.ctab + ((long long)which * av.n_mat1 + m) * nx + col);
        };

        int slot = 0;
        for (long long r = r0; r < r1; ++r) {
            mbar_wait(&bars[slot], (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            const unsigned char *src = ring + slot * kSlotBytes;
            double cur[3][4];
#pragma unroll
            for (int f = 0; f < 3; ++f) {
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

#pragma unroll
            for (int s = 2 * K; s > 0; --s) info[s] = info[s - 1];
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
                // stage s: cur = row q = r - 2s at level s -> cur = row q-2 at level s+1
                const RowInfo &rq = info[2 * s], &r1i = info[2 * s + 1], &r2i = info[2 * s + 2];
                const long long cell_q = cell_r - 2 * s * nx;

                // 1. boundaries and probes of p (row q)
                if (rq.classed & 1u) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        cur[0][c] = apply_class(cls_alpha, cls_value, 0,
                                                (unsigned)(rq.ids >> (16 * c)), cur[0][c]);
                }
                if (rq.flagged) {
                    const long long q = r - 2 * s;
#pragma unroll
                    for (int c = 0; c < 4; ++c) scratch[c] = cur[0][c];
                    stream_slow_cells(a.tables, 0, 1, a.sig_index + s, a.ring_row + s, cell_q, rq.ids,
                                      lane_owned && q >= ys && q < ye, scratch);
#pragma unroll
                    for (int c = 0; c < 4; ++c) cur[0][c] = scratch[c];
                }

                // x-neighbours living in the adjacent lanes (row q-1 operands)
                const double p1_left = shfl_up1(pb[s][0][3]);
                double u1_left = 0, u1_right = 0, v1_left = 0, v1_right = 0;
                if (VISC) {
                    u1_left = shfl_up1(uo[s][0][3]);
                    u1_right = shfl_down1(uo[s][0][0]);
                    v1_left = shfl_up1(vo[s][0][3]);
                    v1_right = shfl_down1(vo[s][0][0]);
                }
                const double un2_right = shfl_down1(un[s][0]);

                // 2. new vx, vy of row q-1; 3. new p of row q-2.
                // coef.*(c): material coefficient of cell c of a row; c = -1 / 4 are the neighbour
                // lanes' cells. Rows: 0 = q, 1 = q-1, 2 = q-2.
                double nu[4], nv[4], np[4];
                auto math = [&](const auto &coef) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double pl = c ? pb[s][0][c - 1] : p1_left;
                        const double du = diff2(coef.gx(1, c - 1), pl, coef.gx(1, c), pb[s][0][c]);
                        const double dv = diff2(coef.gy(2, c), pb[s][1][c], coef.gy(1, c), pb[s][0][c]);
                        const double uold = uo[s][0][c], vold = vo[s][0][c];
                        if (VISC) {
                            const double ul = c ? uo[s][0][c - 1] : u1_left;
                            const double ur = c < 3 ? uo[s][0][c + 1] : u1_right;
                            const double vl = c ? vo[s][0][c - 1] : v1_left;
                            const double vr = c < 3 ? vo[s][0][c + 1] : v1_right;
                            const double cm1 = coef.vm1(c - 1), cp1 = coef.vp1(c + 1);
                            double vis = acc0(mul(coef.vmn(2, c), uo[s][1][c]));
                            vis = add(vis, mul(cm1, ul));
                            vis = add(vis, mul(coef.v0(c), uold));
                            vis = add(vis, mul(cp1, ur));
                            vis = add(vis, mul(coef.vpn(0, c), cur[1][c]));
                            double rhs = sub(du, vis);
                            if (AXI) rhs = add(rhs, mul(coef.eb(c), uold) / coef.rr(c));
                            nu[c] = sub(uold, rhs);
                            double visv = acc0(mul(coef.vmn(2, c), vo[s][1][c]));
                            visv = add(visv, mul(cm1, vl));
                            visv = add(visv, mul(coef.v0(c), vold));
                            visv = add(visv, mul(cp1, vr));
                            visv = add(visv, mul(coef.vpn(0, c), cur[2][c]));
                            nv[c] = sub(vold, sub(dv, visv));
                        } else {
                            nu[c] = AXI ? sub(uold, add(du, mul(0.0, uold))) : sub(uold, du);
                            nv[c] = sub(vold, dv);
                        }
                    }
                    if (r1i.classed & 2u) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            nu[c] = apply_class(cls_alpha, cls_value, 1,
                                                (unsigned)(r1i.ids >> (16 * c)), nu[c]);
                    }
                    if (r1i.classed & 4u) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            nv[c] = apply_class(cls_alpha, cls_value, 2,
                                                (unsigned)(r1i.ids >> (16 * c)), nv[c]);
                    }
                    if (r1i.flagged) {
                        const long long q1 = r - 2 * s - 1;
#pragma unroll
                        for (int c = 0; c < 4; ++c) { scratch[c] = nu[c]; scratch[4 + c] = nv[c]; }
                        stream_slow_cells(a.tables, 1, 2, a.sig_index + s, a.ring_row + s,
                                          cell_q - nx, r1i.ids,
                                          lane_owned && q1 >= ys && q1 < ye, scratch);
#pragma unroll
                        for (int c = 0; c < 4; ++c) { nu[c] = scratch[c]; nv[c] = scratch[4 + c]; }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        double f0 = un[s][c], f1 = c < 3 ? un[s][c + 1] : un2_right;
                        if (AXI) {
                            f0 = mul(f0, coef.r(c));
                            f1 = mul(f1, coef.r(c + 1));
                        }
                        const double divx = diff2(coef.fx(c), f0, coef.fx(c + 1), f1);
                        const double divy = diff2(coef.fy(2, c), vn[s][c], coef.fy(1, c), nv[c]);
                        np[c] = sub(pb[s][1][c], add(divx, divy));
                    }
                };

                const bool uniform = rq.uniform >= 0 && rq.uniform == r1i.uniform &&
                                     rq.uniform == r2i.uniform;
                if (uniform) {
                    const int m = rq.uniform;
                    if (AXI && m != cx_material) {
                        // per-column coefficients of this material for columns x0-1 .. x0+4
#pragma unroll
                        for (int c = 0; c < 5; ++c) cx_fx[c] = ctab(FDS_CTAB_FX, m, cols[c + 1]);
                        if (VISC) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                cx_vm1[c] = ctab(FDS_CTAB_VM1, m, cols[c]);       // column of cell c-1
                                cx_vp1[c] = ctab(FDS_CTAB_VP1, m, cols[c + 2]);   // column of cell c+1
                            }
                        }
                        cx_material = m;
                    }
                    struct {
                        double g0, g1, f0, f1, h0, hp, hm, hmn, hpn, e0;
                        const double *fxc, *vm1c, *vp1c, *rc, *rrc;
                        __device__ double gx(int, int) const { return g0; }
                        __device__ double gy(int, int) const { return g1; }
                        __device__ double fx(int c) const { return AXI ? fxc[c] : f0; }
                        __device__ double fy(int, int) const { return f1; }
                        __device__ double vm1(int c) const { return AXI ? vm1c[c + 1] : hm; }
                        __device__ double vp1(int c) const { return AXI ? vp1c[c - 1] : hp; }
                        __device__ double vmn(int, int) const { return hmn; }
                        __device__ double vpn(int, int) const { return hpn; }
                        __device__ double v0(int) const { return h0; }
                        __device__ double eb(int) const { return e0; }
                        __device__ double r(int c) const { return rc[c]; }
                        __device__ double rr(int c) const { return rrc[c]; }
                    } coef{tabs[FDS_TAB_GX][m], tabs[FDS_TAB_GY][m], tabs[FDS_TAB_FX][m],
                           tabs[FDS_TAB_FY][m], tabs[FDS_TAB_V0][m], tabs[FDS_TAB_VP1][m],
                           tabs[FDS_TAB_VM1][m], tabs[FDS_TAB_VMN][m], tabs[FDS_TAB_VPN][m],
                           tabs[FDS_TAB_EB][m], cx_fx, cx_vm1, cx_vp1, cx_r, cx_rr};
                    math(coef);
                } else {
                    struct {
                        const double (*tabs)[kMaxMaterials];
                        unsigned long long ids[3], left[3], right[3];
                        const double *ctab_base, *rc, *rrc;
                        const long long *cols;
                        long long nx;
                        int n_mat1;
                        // material of cell c (-1 .. 4) of row `row`
                        __device__ int mat(int row, int c) const {
                            if (c < 0) return (int)(left[row] >> 48) & kIdMask;
                            if (c > 3) return (int)right[row] & kIdMask;
                            return (int)(ids[row] >> (16 * c)) & kIdMask;
                        }
                        __device__ double ct(int which, int m, int c) const {   // column of cell c
                            return __ldg(ctab_base + ((long long)which * n_mat1 + m) * nx + cols[c + 1]);
                        }
                        __device__ double gx(int row, int c) const { return tabs[FDS_TAB_GX][mat(row, c)]; }
                        __device__ double gy(int row, int c) const { return tabs[FDS_TAB_GY][mat(row, c)]; }
                        __device__ double fx(int c) const {
                            return AXI ? ct(FDS_CTAB_FX, mat(2, c), c) : tabs[FDS_TAB_FX][mat(2, c)];
                        }
                        __device__ double fy(int row, int c) const { return tabs[FDS_TAB_FY][mat(row, c)]; }
                        __device__ double vm1(int c) const {
                            return AXI ? ct(FDS_CTAB_VM1, mat(1, c), c) : tabs[FDS_TAB_VM1][mat(1, c)];
                        }
                        __device__ double vp1(int c) const {
                            return AXI ? ct(FDS_CTAB_VP1, mat(1, c), c) : tabs[FDS_TAB_VP1][mat(1, c)];
                        }
                        __device__ double vmn(int row, int c) const { return tabs[FDS_TAB_VMN][mat(row, c)]; }
                        __device__ double vpn(int row, int c) const { return tabs[FDS_TAB_VPN][mat(row, c)]; }
                        __device__ double v0(int c) const { return tabs[FDS_TAB_V0][mat(1, c)]; }
                        __device__ double eb(int c) const { return tabs[FDS_TAB_EB][mat(1, c)]; }
                        __device__ double r(int c) const { return rc[c]; }
                        __device__ double rr(int c) const { return rrc[c]; }
                    } coef{tabs,
                           {rq.ids, r1i.ids, r2i.ids},
                           {0ull, __shfl_up_sync(0xffffffffu, r1i.ids, 1), 0ull},
                           {0ull, __shfl_down_sync(0xffffffffu, r1i.ids, 1),
                            __shfl_down_sync(0xffffffffu, r2i.ids, 1)},
                           av.ctab, cx_r, cx_rr, cols, nx, av.n_mat1};
                    math(coef);
                }

                // 4. row q-2 at level s+1 goes to the next stage; shift the window
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double keep_p = cur[0][c], keep_u = cur[1][c], keep_v = cur[2][c];
                    cur[0][c] = np[c];
                    cur[1][c] = un[s][c];
                    cur[2][c] = vn[s][c];
                    pb[s][1][c] = pb[s][0][c];
                    pb[s][0][c] = keep_p;
                    uo[s][1][c] = uo[s][0][c];
                    uo[s][0][c] = keep_u;
                    vo[s][1][c] = vo[s][0][c];
                    vo[s][0][c] = keep_v;
                    un[s][c] = nu[c];
                    vn[s][c] = nv[c];
                }
            }

            const long long orow = r - kLag;
            if (lane_owned && orow >= ys && orow < ye) {
                const long long o = cell_r - kLag * nx;
#pragma unroll
                for (int f = 0; f < 3; ++f) {
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
