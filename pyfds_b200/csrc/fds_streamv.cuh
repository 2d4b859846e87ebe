// Streaming kernel for the viscous (lossy) Acoustic2D leapfrog and for lossy Acoustic3DAxi:
// pyfds/acoustics.py:111-128 and 205-225, K <= 2 time steps per launch. (The lossless axisymmetric
// model runs on stream2d_kernel<AXI>; the VISC = false instantiation of this template is not
// dispatched.)
//
// Same machinery and geometry as stream2d_kernel (fds_stream2d.cuh): one warp = one 64-cell strip
// (2 cells per lane, 4 halo cells either side, 56 owned) x a chunk of rows, rows arriving in pairs
// through the bulk-copy ring, a K-stage time pipeline in registers, x-neighbours by warp shuffle,
// tasks from the cost-balanced table. What differs is the stage: the viscous 5-point operator works
// on the OLD velocities, so a stage needs a three-row window and lags TWO rows behind its input:
//
//   row q arrives at level s  ->  new vx, vy of row q-1 (needs old vx, vy of rows q-2, q-1, q and p of
//   rows q-2, q-1)            ->  new p of row q-2 (needs new vx of row q-2, new vy of rows q-2, q-1)
//
// Per stage eight row fragments stay in registers (p after boundaries, old vx, old vy of rows q-1 and
// q-2; new vx, vy of row q-2): 32 registers per lane and stage. The dependency cone grows one cell to
// the left and two to the right per step, which the 4-cell strip halo covers for K <= 2.
//
// STEADY rows (no table lookup, no probe, the lane's map word repeating row after row; one material
// with class operations on at most one component, or several materials -- interfaces along y --
// without operations) run through a branch-free body that takes two rows per
// iteration; the two window rows swap roles between the halves of the iteration, so no register is
// moved. Material coefficients are loaded into registers when such a run starts; for the
// axisymmetric model the per-column values of this lane's cells and of its two neighbours (a_p_vx / r,
// the viscous x diagonals with the 1/r term, r, r^2, 1 / r^2) go to per-lane shared-memory slots and
// are re-read for every row (FDS_SV_COLSMEM). Every other row takes the general row iteration: one
// rolled copy of the stage code with per-cell coefficient lookups, inline classes and the table slow
// path.
//
// Every value is produced by the same __dmul_rn/__dadd_rn sequence as cell_body in fds_step2d.cuh;
// the one exception, the axisymmetric model's division by r^2 in steady rows, is a sequence that
// returns the IEEE quotient bit for bit (sv_quotient below).
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

#ifndef FDS_SV_CTAS
#define FDS_SV_CTAS (12 / FDS_STREAM_WARPS)
#endif
constexpr int kSVCtasPerSm = FDS_SV_CTAS;                   // resident CTAs per SM (register budget: 168)
constexpr int kMaxStreamVSteps = 2;         // K: bounded by the strip halo (see above)

// Coefficients of this lane's cells on steady rows (the same for rows q, q-1, q-2: steady rows repeat
// their map words). Rows of one material fill every entry from the same register; members a model
// does not use are never loaded (the struct lives in registers, everything is inlined).
struct SVCoef {
    static constexpr int C = kS2LaneCells;
    double gx[C + 1];            // a_vx_p factor of cell c-1 (gx[0]: the left-hand lane's last cell)
    double gy[C], fy[C];         // a_vy_p, a_p_vy factors of cell c
    double fx[C + 1];            // a_p_vx factor of cell c (fx[C]: the right-hand lane's first cell);
                                 // axisymmetric: divided by r of the column
    double v0[C], vmn[C], vpn[C];   // a_v_v main and +-nx diagonals of cell c
    double vm1[C], vp1[C];       // a_v_v x diagonals: entry of cell c-1 / of cell c+1
    double eb[C];                // axisymmetric extra term dt*mu/rho of cell c
    double r[C + 1], rr[C];      // axisymmetric: r of columns x0 .. x0+C, r^2 of x0 .. x0+C-1
    double ry[C];                // axisymmetric: 1 / r^2 rounded to nearest (sv_quotient)
};

// The axisymmetric model divides by r^2 in every cell and stage (pyfds/acoustics.py:210-212). nvcc
// expands an IEEE division into a reciprocal seed, four Newton steps, the quotient, a residual
// correction and a guarded call of a slow path: ~22 instructions in one serial chain, and the guard's
// branch cuts the stage into basic blocks the scheduler cannot move work across. The divisor only
// depends on the column, so its correctly rounded reciprocal y = RN(1 / b) is computed once per steady
// run (with the IEEE division) and a cell needs three operations:
//     q0 = RN(a y);   t = b q0 - a  (exact, one fma);   q = RN(q0 - t y)
// q is the correctly rounded a / b (Markstein 1990: y correctly rounded, q0 within one ulp, the
// residual exact) as long as nothing under- or overflows on the way and the significand of b is not
// all ones: 2^-256 <= b <= 2^256 and 2^-693 <= |a| <= 2^677 guarantee that (the residual is a
// multiple of 2^(e_a - 106) or larger, q0 and q stay normal). Written as q0 - t y, signed zeros come
// out right as well: a = +-0 gives q0 = +-0, t = +0, q = +-0. The numerator is a = eb * vx with eb a
// material constant, so "vx is zero or 2^-493 <= |vx| <= 2^477" (sv_covered) together with
// 2^-200 <= eb <= 2^200 is sufficient; the steady loop checks the four rows of vx an iteration
// divides BEFORE it starts the iteration and runs the iteration with the IEEE division otherwise
// (denormals and the leading edge of a pulse that decays towards them, infinities, NaN).
// tests/test_fast_division.py runs the same sequence against the division on the CPU.
constexpr unsigned kDivHiLo = (1023u - 493u) << 20;    // high word of 2^-493
constexpr unsigned kDivHiSpan = (493u + 477u) << 20;   // ... up to 2^477
__device__ __forceinline__ double sv_quotient(double a, double b, double y) {
    const double q0 = mul(a, y);
    const double t = __fma_rn(b, q0, -a);
    return __fma_rn(-t, y, q0);
}
__device__ __forceinline__ bool sv_covered(double vx) {
    const unsigned h = (unsigned)__double2hiint(vx) & 0x7fffffffu;
    return h - kDivHiLo < kDivHiSpan || (h | (unsigned)__double2loint(vx)) == 0u;
}
// positive, exponent within +-`range`, significand not all ones
__device__ __forceinline__ bool sv_moderate(double b, unsigned range) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(b);
    const unsigned e = (unsigned)(u >> 52);
    return e >= 1023u - range && e <= 1023u + range &&
           (u & 0xfffffffffffffull) != 0xfffffffffffffull;
}
// Build switches for A/B measurements (python -m pyfds_b200._build --variant NAME -DFDS_SV_...=0):
// 0 = the IEEE division in every cell / the per-column coefficients in registers, as before round 2.
#ifndef FDS_SV_FASTDIV
#define FDS_SV_FASTDIV 1
#endif
#ifndef FDS_SV_COLSMEM
#define FDS_SV_COLSMEM 1
#endif

// One stage on a steady row: `cur` = row q at level s on entry, row q-2 at level s+1 on exit.
// pA/uA/vA = p (after boundaries), old vx, old vy of row q-1; pB/uB/vB = the same of row q-2 on entry
// and of row q on exit (the next row's "A": the caller swaps the roles). U/V = new vx, vy of row q-2
// (read), Un/Vn = new vx, vy of row q-1 (written). CC, ca, cv: as in steady_stage (fds_stream2d.cuh).
template <bool AXI, bool VISC, int CC, bool EXACT = true>
__device__ __forceinline__ void sv_steady_stage(
    double (&cur)[3][kS2LaneCells], const double (&pA)[kS2LaneCells],
    const double (&uA)[kS2LaneCells], const double (&vA)[kS2LaneCells], double (&pB)[kS2LaneCells],
    double (&uB)[kS2LaneCells], double (&vB)[kS2LaneCells], const double (&U)[kS2LaneCells],
    const double (&V)[kS2LaneCells], double (&Un)[kS2LaneCells], double (&Vn)[kS2LaneCells],
    const SVCoef &ku, const double (&ca)[kS2LaneCells], const double (&cv)[kS2LaneCells],
    const double2 *col = nullptr) {
    constexpr int C = kS2LaneCells;
    static_assert(C == 2, "the per-column slots are pairs");
    SVCoef k = ku;
    if (AXI && FDS_SV_COLSMEM) {
        // per-column coefficients of this lane: re-read for every stage instead of held in registers
        const double2 a0 = col[0 * 32], a1 = col[1 * 32], a2 = col[2 * 32], a3 = col[3 * 32];
        const double2 a4 = col[4 * 32], a5 = col[5 * 32], a6 = col[6 * 32];
        k.fx[0] = a0.x; k.fx[1] = a0.y; k.fx[2] = a1.x; k.r[0] = a1.y; k.r[1] = a2.x; k.r[2] = a2.y;
        k.vm1[0] = a3.x; k.vm1[1] = a3.y; k.vp1[0] = a4.x; k.vp1[1] = a4.y;
        k.rr[0] = a5.x; k.rr[1] = a5.y; k.ry[0] = a6.x; k.ry[1] = a6.y;
    }
    if (CC == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) cur[0][c] = add(mul(ca[c], cur[0][c]), cv[c]);
    }
    const double p1_left = shfl_up1(pA[C - 1]);
    double u1_left = 0, u1_right = 0, v1_left = 0, v1_right = 0;
    if (VISC) {
        u1_left = shfl_up1(uA[C - 1]);
        u1_right = shfl_down1(uA[0]);
        v1_left = shfl_up1(vA[C - 1]);
        v1_right = shfl_down1(vA[0]);
    }
    const double un2_right = shfl_down1(U[0]);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double pl = c ? pA[c - 1] : p1_left;
        const double du = diff2(k.gx[c], pl, k.gx[c + 1], pA[c]);
        const double dv = diff2(k.gy[c], pB[c], k.gy[c], pA[c]);
        const double uold = uA[c], vold = vA[c];
        if (VISC) {
            const double ul = c ? uA[c - 1] : u1_left;
            const double ur = c < C - 1 ? uA[c + 1] : u1_right;
            const double vl = c ? vA[c - 1] : v1_left;
            const double vr = c < C - 1 ? vA[c + 1] : v1_right;
            double vis = acc0(mul(k.vmn[c], uB[c]));
            vis = add(vis, mul(k.vm1[c], ul));
            vis = add(vis, mul(k.v0[c], uold));
            vis = add(vis, mul(k.vp1[c], ur));
            vis = add(vis, mul(k.vpn[c], cur[1][c]));
            double rhs = sub(du, vis);
            if (AXI)
                rhs = add(rhs, EXACT ? mul(k.eb[c], uold) / k.rr[c]
                                     : sv_quotient(mul(k.eb[c], uold), k.rr[c], k.ry[c]));
            Un[c] = sub(uold, rhs);
            double visv = acc0(mul(k.vmn[c], vB[c]));
            visv = add(visv, mul(k.vm1[c], vl));
            visv = add(visv, mul(k.v0[c], vold));
            visv = add(visv, mul(k.vp1[c], vr));
            visv = add(visv, mul(k.vpn[c], cur[2][c]));
            Vn[c] = sub(vold, sub(dv, visv));
        } else {
            Un[c] = AXI ? sub(uold, add(du, mul(0.0, uold))) : sub(uold, du);
            Vn[c] = sub(vold, dv);
        }
        if (CC == 1) Un[c] = add(mul(ca[c], Un[c]), cv[c]);
        if (CC == 2) Vn[c] = add(mul(ca[c], Vn[c]), cv[c]);
    }
    double np[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        double f0 = U[c], f1 = c < C - 1 ? U[c + 1] : un2_right;
        if (AXI) {
            f0 = mul(f0, k.r[c]);
            f1 = mul(f1, k.r[c + 1]);
        }
        const double divx = diff2(k.fx[c], f0, k.fx[c + 1], f1);
        const double divy = diff2(k.fy[c], V[c], k.fy[c], Vn[c]);
        np[c] = sub(pB[c], add(divx, divy));
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        pB[c] = cur[0][c];
        uB[c] = cur[1][c];
        vB[c] = cur[2][c];
        cur[0][c] = np[c];
        cur[1][c] = U[c];
        cur[2][c] = V[c];
    }
}

template <int K, bool AXI, bool VISC, int CTAS = kSVCtasPerSm>
__global__ void __launch_bounds__(kStreamWarps * 32, CTAS) streamv_kernel(StreamVArgs av) {
    constexpr int C = kS2LaneCells;
    constexpr int kLag = 2 * K;     // rows between the input row and the row stored
    constexpr int W = 2 * K + 1;    // row tags in flight: rows r .. r-2K
    static_assert(K >= 1 && K <= kMaxStreamVSteps, "the strip halo covers two steps");
    const Stream2DArgs &a = av.base;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double tabs[FDS_TAB_COUNT][kMaxMaterials];
    __shared__ double cls_alpha[3][kMaxClasses], cls_value[K][3][kMaxClasses];
    __shared__ double2 colco[AXI && FDS_SV_COLSMEM ? kStreamWarps * 7 * 32 : 1];

    for (int k = threadIdx.x; k < FDS_TAB_COUNT * kMaxMaterials; k += blockDim.x)
        (&tabs[0][0])[k] = a.tab[k];
    stage_class_tables<K>(a, cls_alpha, cls_value);
    __syncthreads();
    // sweep overlap: see TaskSync in fds_stream2d.cuh
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nx = a.nx;
    unsigned char *ring = smem_raw + warp * kS2WarpRingBytes;
    double *scratch =
        reinterpret_cast<double *>(ring + kS2RingDepth * kS2SlotBytes) + lane * 2 * C;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        ring + kS2RingDepth * kS2SlotBytes + kS2ScratchBytes);
    int *stash = reinterpret_cast<int *>(bars + kS2RingDepth / 2);
    unsigned phase_bits = 0;   // parity of every pair barrier (the barriers live across tasks)

    if (lane == 0) {
        for (int d = 0; d < kS2RingDepth / 2; ++d) mbar_init(&bars[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(a.task_counter, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const int4 tk = __ldg(a.tasks + task);
        const int ys = tk.y, ye = tk.z;
        const long long xs = (long long)tk.x * kS2StripStride - kS2StripHalo;   // column of lane 0
        const int r0 = ys - kLag, r1 = ye + kLag;                               // rows streamed in
        if (lane == 0) {
            stash[0] = task;
            stash[1] = tk.w;
        }
        task_acquire(a.sync, task, tk.w, lane);

        // rows r, r+1 (r - r0 even) into ring slots `slot`, `slot` + 1 (slot even), one barrier
        auto issue_pair = [&](int r, int slot) {
            const int n_rows = r + 1 < r1 ? 2 : 1;
            void *bar = &bars[slot >> 1];
            mbar_expect_tx(bar, n_rows * (3 * kS2FieldBytes + kS2MapWindowBytes));
            long long base = (long long)r * nx + xs;   // flat cell index of the strip start
            unsigned char *dst = ring + slot * kS2SlotBytes;
            for (int h = 0; h < n_rows; ++h) {
                bulk_load(dst, a.in[0] + base, kS2FieldBytes, bar);
                bulk_load(dst + kS2FieldBytes, a.in[1] + base, kS2FieldBytes, bar);
                bulk_load(dst + 2 * kS2FieldBytes, a.in[2] + base, kS2FieldBytes, bar);
                bulk_load(dst + 3 * kS2FieldBytes, a.map + (base & ~7LL), kS2MapWindowBytes, bar);
                base += nx;
                dst += kS2SlotBytes;
            }
        };
        if (lane == 0)
            for (int d = 0; d < kS2RingDepth && r0 + d < r1; d += 2) issue_pair(r0 + d, d);
        // optional counters (tests): as in stream2d_kernel
        if (a.stats && lane == 0) atomicAdd(a.stats + 6, (unsigned long long)(r1 - r0));

        // pipeline state per stage: p after boundaries, old vx, old vy of rows q-1 (A) and q-2 (B);
        // new vx, vy of row q-2
        double pA[K][C], uA[K][C], vA[K][C], pB[K][C], uB[K][C], vB[K][C], un[K][C], vn[K][C];
        RowTag info[W];
#pragma unroll
        for (int s = 0; s < K; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                pA[s][c] = uA[s][c] = vA[s][c] = pB[s][c] = uB[s][c] = vB[s][c] = 0.0;
                un[s][c] = vn[s][c] = 0.0;
            }
#pragma unroll
        for (int s = 0; s < W; ++s) info[s] = RowTag{0u, 0u};
        int run = 0;

        const long long x0 = xs + C * lane;
        const bool lane_owned = lane >= kS2HaloLanes && lane < 32 - kS2HaloLanes && x0 < nx;
        const bool lane_relevant = x0 < nx + kS2StripHalo;
        long long cell_r = (long long)r0 * nx + x0;   // flat index of this lane's first cell in row r
        const int map_step = (int)(nx & 7LL);
        int map_off = (int)(((long long)r0 * nx + xs) & 7LL);   // entry offset in the map window

        // column of cell c (-1 .. C) of this lane, wrapped into the row like the flat index
        auto col_of = [&](int c) {
            long long col = x0 + c;
            if (col < 0) col += nx;
            if (col >= nx) col -= nx;
            if (col >= nx) col -= nx;
            return col;
        };
        auto ctab = [&](int which, int m, long long col) {
            return __ldg(av.ctab + ((long long)which * av.n_mat1 + m) * nx + col);
        };

        auto load_row = [&](double (&cur)[3][C], const unsigned char *src) {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const double2 v =
                    *reinterpret_cast<const double2 *>(src + f * kS2FieldBytes + lane * 16);
                cur[f][0] = v.x;
                cur[f][1] = v.y;
            }
        };
        auto store_row = [&](const double (&cur)[3][C], int orow, long long o) {
            if (lane_owned && orow >= ys && orow < ye) {
#pragma unroll
                for (int f = 0; f < 3; ++f)
                    __stcs(reinterpret_cast<double2 *>(a.out[f] + o),
                           make_double2(cur[f][0], cur[f][1]));
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
            RowTag m = RowTag{0u, 32u};   // past the last row: never steady
            if (fetch_row < r1) {
                await(fetch_row, fetch_slot);
                m = s2_row_meta(ring + fetch_slot * kS2SlotBytes, map_off, lane, lane_relevant);
                map_off = (map_off + map_step) & 7;
                if (++fetch_slot == kS2RingDepth) fetch_slot = 0;
                ++fetch_row;
            }
            return m;
        };
        // (the axisymmetric model keeps rows of several materials on the general row iteration: per-lane
        // copies of its per-column AND per-material coefficients do not fit the register budget)
        auto sv_steady = [](const RowTag &t) { return t.steady() && !(AXI && (t.bits & 32u)); };
        RowTag m0 = fetch(), m1 = fetch();
        // What the pipeline computes from the rows above r0 (zero state) never reaches an owned row,
        // so those rows may as well count as rows like the first one (see stream2d_kernel).
        if (sv_steady(m0)) {
#pragma unroll
            for (int s = 0; s < W; ++s) info[s] = m0;
            run = W;
        }

        int slot = 0;
        int r = r0;
        // ---- steady pairs: rows r-2K .. r+1 carry the same map words (one material, no table
        // lookup, no probe, constant operations on component CC only, or none: CC = -1) -------------
        auto steady_pairs = [&](auto cc_tag, auto uni_tag) {
            constexpr int CC = decltype(cc_tag)::value;
            constexpr bool UNI = decltype(uni_tag)::value;   // one material: warp-uniform coefficients
            if (a.stats && lane == 0) atomicAdd(a.stats + (UNI ? CC + 1 : 4), 1ull);
            const unsigned my_ids = info[0].ids;
            SVCoef k;
            constexpr bool kFast = AXI && VISC && FDS_SV_FASTDIV;   // sv_quotient
            bool fast_run = true;   // this lane's r^2 and eb are in the range sv_quotient covers
            {
                const unsigned material = info[0].bits & kIdMask;
                const unsigned left = __shfl_up_sync(0xffffffffu, my_ids, 1) >> (16 * (C - 1));
                const unsigned right = __shfl_down_sync(0xffffffffu, my_ids, 1);
                // material of cell c (-1 .. C) of this lane
                auto mat = [&](int c) {
                    if (UNI) return material;
                    return (c < 0 ? left : c >= C ? right : my_ids >> (16 * c)) & kIdMask;
                };
                // one material: every entry of a table comes from ONE register
                auto tab = [&](int which, int c) {
                    return tabs[which][mat(c)];
                };
                const double u_gx = tabs[FDS_TAB_GX][material], u_gy = tabs[FDS_TAB_GY][material];
                const double u_fx = tabs[FDS_TAB_FX][material], u_fy = tabs[FDS_TAB_FY][material];
                const double u_v0 = tabs[FDS_TAB_V0][material], u_eb = tabs[FDS_TAB_EB][material];
                const double u_vmn = tabs[FDS_TAB_VMN][material];
                const double u_vpn = tabs[FDS_TAB_VPN][material];
                const double u_vm1 = tabs[FDS_TAB_VM1][material];
                const double u_vp1 = tabs[FDS_TAB_VP1][material];
#pragma unroll
                for (int c = 0; c <= C; ++c) {
                    k.gx[c] = UNI ? u_gx : tab(FDS_TAB_GX, c - 1);
                    k.fx[c] = AXI ? ctab(FDS_CTAB_FX, mat(c), col_of(c))
                                  : UNI ? u_fx : tab(FDS_TAB_FX, c);
                    if (AXI) k.r[c] = __ldg(av.cvec + FDS_CVEC_R * nx + col_of(c));
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    k.gy[c] = UNI ? u_gy : tab(FDS_TAB_GY, c);
                    k.fy[c] = UNI ? u_fy : tab(FDS_TAB_FY, c);
                    if (AXI) k.rr[c] = __ldg(av.cvec + FDS_CVEC_RR * nx + col_of(c));
                    k.ry[c] = 0.0;
                    if (kFast) {
                        fast_run = fast_run && sv_moderate(k.rr[c], 256u);
                        k.ry[c] = 1.0 / k.rr[c];
                    }
                    if (VISC) {
                        k.v0[c] = UNI ? u_v0 : tab(FDS_TAB_V0, c);
                        k.vmn[c] = UNI ? u_vmn : tab(FDS_TAB_VMN, c);
                        k.vpn[c] = UNI ? u_vpn : tab(FDS_TAB_VPN, c);
                        k.eb[c] = UNI ? u_eb : tab(FDS_TAB_EB, c);
                        k.vm1[c] = AXI ? ctab(FDS_CTAB_VM1, mat(c - 1), col_of(c - 1))
                                       : UNI ? u_vm1 : tab(FDS_TAB_VM1, c - 1);
                        k.vp1[c] = AXI ? ctab(FDS_CTAB_VP1, mat(c + 1), col_of(c + 1))
                                       : UNI ? u_vp1 : tab(FDS_TAB_VP1, c + 1);
                        if (kFast) fast_run = fast_run && (k.eb[c] == 0.0 || sv_moderate(k.eb[c], 200u));
                    }
                }
            }
            double2 *col = colco + (AXI && FDS_SV_COLSMEM ? warp * 7 * 32 + lane : 0);
            if (AXI && FDS_SV_COLSMEM) {
                __syncwarp();
                col[0 * 32] = make_double2(k.fx[0], k.fx[1]);
                col[1 * 32] = make_double2(k.fx[2], k.r[0]);
                col[2 * 32] = make_double2(k.r[1], k.r[2]);
                col[3 * 32] = make_double2(k.vm1[0], k.vm1[1]);
                col[4 * 32] = make_double2(k.vp1[0], k.vp1[1]);
                col[5 * 32] = make_double2(k.rr[0], k.rr[1]);
                col[6 * 32] = make_double2(k.ry[0], k.ry[1]);
                __syncwarp();
            }
            double ca[C], cv[K][C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned kc =
                    CC < 0 ? 0u : (my_ids >> (16 * c + class_shift(CC < 0 ? 0 : CC))) & 7u;
                ca[c] = kc ? cls_alpha[CC < 0 ? 0 : CC][kc] : 1.0;
#pragma unroll
                for (int s = 0; s < K; ++s) cv[s][c] = kc ? cls_value[s][CC < 0 ? 0 : CC][kc] : -0.0;
            }
            // rows r and r+1 through the K stages (`exact_tag`: with the IEEE division)
            auto two_rows = [&](double (&cur)[3][C], auto exact_tag) {
                constexpr bool EXACT = decltype(exact_tag)::value;
                double un1[K][C], vn1[K][C];
#pragma unroll
                for (int s = 0; s < K; ++s)
                    sv_steady_stage<AXI, VISC, CC, EXACT>(cur, pA[s], uA[s], vA[s], pB[s], uB[s],
                                                          vB[s], un[s], vn[s], un1[s], vn1[s], k, ca,
                                                          cv[s], col);
                store_row(cur, r - kLag, cell_r - kLag * nx);

                load_row(cur, ring + (slot + 1) * kS2SlotBytes);
                __syncwarp();
                if (lane == 0 && r + kS2RingDepth < r1) issue_pair(r + kS2RingDepth, slot);
#pragma unroll
                for (int s = 0; s < K; ++s)
                    sv_steady_stage<AXI, VISC, CC, EXACT>(cur, pB[s], uB[s], vB[s], pA[s], uA[s],
                                                          vA[s], un1[s], vn1[s], un[s], vn[s], k, ca,
                                                          cv[s], col);
                store_row(cur, r + 1 - kLag, cell_r + nx - kLag * nx);
            };
            for (;;) {
                double cur[3][C];
                load_row(cur, ring + slot * kS2SlotBytes);
                if (kFast) {
                    // the vx rows this iteration divides: stage s of row r takes the old vx of its
                    // row q-1 (uA[s]), stage s of row r+1 what stage s of row r receives as row q
                    // (the row just loaded; the new vx the stage before has in hand)
                    bool covered = true;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        covered &= sv_covered(cur[1][c]);
#pragma unroll
                        for (int s = 0; s < K; ++s) {
                            covered &= sv_covered(uA[s][c]);
                            if (s + 1 < K) covered &= sv_covered(un[s][c]);
                        }
                    }
                    if (__builtin_expect(
                            __all_sync(0xffffffffu, !lane_relevant || (covered && fast_run)), 1)) {
                        two_rows(cur, std::false_type{});
                    } else {
                        if (a.stats && lane == 0) atomicAdd(a.stats + 7, 1ull);
                        two_rows(cur, std::true_type{});
                    }
                } else {
                    two_rows(cur, std::true_type{});
                }
                cell_r += 2 * nx;
                r += 2;
                slot = slot + 2 == kS2RingDepth ? 0 : slot + 2;
                // the lookahead is implied in here: rows r, r+1 have landed and been examined
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
            if (run >= W && !(slot & 1) && sv_steady(m0) && m0.bits == info[0].bits &&
                m1.bits == m0.bits &&
                (m0.plain() ||
                 __all_sync(0xffffffffu, m0.ids == info[0].ids && m1.ids == info[0].ids))) {
                using std::integral_constant;
                using std::true_type;
                if (!AXI && (m0.bits & 32u))
                    steady_pairs(integral_constant<int, -1>{}, std::false_type{});
                else
                    switch (m0.classed()) {
                        case 0u: steady_pairs(integral_constant<int, -1>{}, true_type{}); break;
                        case 1u: steady_pairs(integral_constant<int, 0>{}, true_type{}); break;
                        case 2u: steady_pairs(integral_constant<int, 1>{}, true_type{}); break;
                        default: steady_pairs(integral_constant<int, 2>{}, true_type{}); break;
                    }
                continue;
            }

            // ---- general row iteration -----------------------------------------------------------
            if (a.stats && lane == 0) atomicAdd(a.stats + 5, 1ull);
            double cur[3][C];
            load_row(cur, ring + slot * kS2SlotBytes);
            __syncwarp();
            // the pair of ring slots is free once its second row has been read
            if (lane == 0 && (slot & 1) && r - 1 + kS2RingDepth < r1)
                issue_pair(r - 1 + kS2RingDepth, slot - 1);

#pragma unroll
            for (int s = W - 1; s > 0; --s) info[s] = info[s - 1];
            info[0] = m0;
            // run = rows in a row, the last one being info[0], that are steady with the same map words
            if (!sv_steady(info[0]))
                run = 0;
            else if (run > 0 && info[0].bits == info[1].bits &&
                     (info[0].plain() || __all_sync(0xffffffffu, info[0].ids == info[1].ids)))
                ++run;
            else
                run = 1;

            // The K stages run as a rolled loop over one copy of the stage code: stage s works on
            // state index 0 and on the tags info[0..2] (rows q, q-1, q-2); after each stage the state
            // rotates by one place and the tags by two. K rotations put the state back in place, the
            // 2K + 1 tags need one more place after the loop.
#pragma unroll 1
            for (int s = 0; s < K; ++s) {
                // stage s: cur = row q = r - 2s at level s  ->  cur = row q-2 at level s+1
                const RowTag rq = info[0], r1i = info[1], r2i = info[W > 2 ? 2 : 0];
                const int q = r - 2 * s;
                const long long cell_q = cell_r - 2 * s * nx;

                // 1. boundaries and probes of p (row q)
                if (rq.classed() & 1u) {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        cur[0][c] =
                            apply_class(cls_alpha, cls_value[s], 0, rq.ids >> (16 * c), cur[0][c]);
                }
                if (rq.flagged()) {
#pragma unroll
                    for (int c = 0; c < C; ++c) scratch[c] = cur[0][c];
                    s2_slow_cells(a.tables, 0, 1, a.sig_index + s, a.ring_row + s, cell_q, rq.ids,
                                  lane_owned && q >= ys && q < ye, scratch);
#pragma unroll
                    for (int c = 0; c < C; ++c) cur[0][c] = scratch[c];
                }

                // material of cell c (-1 .. C) of row 0 = q, 1 = q-1, 2 = q-2
                const unsigned left1 = __shfl_up_sync(0xffffffffu, r1i.ids, 1);
                const unsigned right1 = __shfl_down_sync(0xffffffffu, r1i.ids, 1);
                const unsigned right2 = __shfl_down_sync(0xffffffffu, r2i.ids, 1);
                auto mat = [&](int row, int c) {
                    if (c < 0) return (int)((left1 >> (16 * (C - 1))) & kIdMask);     // row 1 only
                    if (c >= C) return (int)((row == 1 ? right1 : right2) & kIdMask);
                    const unsigned ids = row == 0 ? rq.ids : row == 1 ? r1i.ids : r2i.ids;
                    return (int)((ids >> (16 * c)) & kIdMask);
                };
                auto vm1_of = [&](int c) {   // x diagonal at offset -1: coefficient of cell c (row q-1)
                    return AXI ? ctab(FDS_CTAB_VM1, mat(1, c), col_of(c)) : tabs[FDS_TAB_VM1][mat(1, c)];
                };
                auto vp1_of = [&](int c) {
                    return AXI ? ctab(FDS_CTAB_VP1, mat(1, c), col_of(c)) : tabs[FDS_TAB_VP1][mat(1, c)];
                };
                auto fx_of = [&](int c) {    // row q-2
                    return AXI ? ctab(FDS_CTAB_FX, mat(2, c), col_of(c)) : tabs[FDS_TAB_FX][mat(2, c)];
                };

                // x-neighbours living in the adjacent lanes (row q-1 operands)
                const double p1_left = shfl_up1(pA[0][C - 1]);
                double u1_left = 0, u1_right = 0, v1_left = 0, v1_right = 0;
                if (VISC) {
                    u1_left = shfl_up1(uA[0][C - 1]);
                    u1_right = shfl_down1(uA[0][0]);
                    v1_left = shfl_up1(vA[0][C - 1]);
                    v1_right = shfl_down1(vA[0][0]);
                }
                const double un2_right = shfl_down1(un[0][0]);

                // 2. new vx, vy of row q-1
                double nu[C], nv[C], np[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double pl = c ? pA[0][c - 1] : p1_left;
                    const double du = diff2(tabs[FDS_TAB_GX][mat(1, c - 1)], pl,
                                            tabs[FDS_TAB_GX][mat(1, c)], pA[0][c]);
                    const double dv = diff2(tabs[FDS_TAB_GY][mat(2, c)], pB[0][c],
                                            tabs[FDS_TAB_GY][mat(1, c)], pA[0][c]);
                    const double uold = uA[0][c], vold = vA[0][c];
                    if (VISC) {
                        const double ul = c ? uA[0][c - 1] : u1_left;
                        const double ur = c < C - 1 ? uA[0][c + 1] : u1_right;
                        const double vl = c ? vA[0][c - 1] : v1_left;
                        const double vr = c < C - 1 ? vA[0][c + 1] : v1_right;
                        const double cm1 = vm1_of(c - 1), cp1 = vp1_of(c + 1);
                        const double c0 = tabs[FDS_TAB_V0][mat(1, c)];
                        const double cmn = tabs[FDS_TAB_VMN][mat(2, c)];
                        const double cpn = tabs[FDS_TAB_VPN][mat(0, c)];
                        double vis = acc0(mul(cmn, uB[0][c]));
                        vis = add(vis, mul(cm1, ul));
                        vis = add(vis, mul(c0, uold));
                        vis = add(vis, mul(cp1, ur));
                        vis = add(vis, mul(cpn, cur[1][c]));
                        double rhs = sub(du, vis);
                        if (AXI)
                            rhs = add(rhs, mul(tabs[FDS_TAB_EB][mat(1, c)], uold) /
                                               __ldg(av.cvec + FDS_CVEC_RR * nx + col_of(c)));
                        nu[c] = sub(uold, rhs);
                        double visv = acc0(mul(cmn, vB[0][c]));
                        visv = add(visv, mul(cm1, vl));
                        visv = add(visv, mul(c0, vold));
                        visv = add(visv, mul(cp1, vr));
                        visv = add(visv, mul(cpn, cur[2][c]));
                        nv[c] = sub(vold, sub(dv, visv));
                    } else {
                        nu[c] = AXI ? sub(uold, add(du, mul(0.0, uold))) : sub(uold, du);
                        nv[c] = sub(vold, dv);
                    }
                }
                if (r1i.classed() & 2u) {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        nu[c] = apply_class(cls_alpha, cls_value[s], 1, r1i.ids >> (16 * c), nu[c]);
                }
                if (r1i.classed() & 4u) {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        nv[c] = apply_class(cls_alpha, cls_value[s], 2, r1i.ids >> (16 * c), nv[c]);
                }
                if (r1i.flagged()) {
                    const int q1 = q - 1;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        scratch[c] = nu[c];
                        scratch[C + c] = nv[c];
                    }
                    s2_slow_cells(a.tables, 1, 2, a.sig_index + s, a.ring_row + s, cell_q - nx,
                                  r1i.ids, lane_owned && q1 >= ys && q1 < ye, scratch);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        nu[c] = scratch[c];
                        nv[c] = scratch[C + c];
                    }
                }
                // 3. new p of row q-2
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    double f0 = un[0][c], f1 = c < C - 1 ? un[0][c + 1] : un2_right;
                    if (AXI) {
                        f0 = mul(f0, __ldg(av.cvec + FDS_CVEC_R * nx + col_of(c)));
                        f1 = mul(f1, __ldg(av.cvec + FDS_CVEC_R * nx + col_of(c + 1)));
                    }
                    const double divx = diff2(fx_of(c), f0, fx_of(c + 1), f1);
                    const double divy = diff2(tabs[FDS_TAB_FY][mat(2, c)], vn[0][c],
                                              tabs[FDS_TAB_FY][mat(1, c)], nv[c]);
                    np[c] = sub(pB[0][c], add(divx, divy));
                }

                // 4. row q-2 at level s+1 goes to the next stage; shift the window; rotate
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double keep_p = cur[0][c], keep_u = cur[1][c], keep_v = cur[2][c];
                    const double old_pa = pA[0][c], old_ua = uA[0][c], old_va = vA[0][c];
                    cur[0][c] = np[c];
                    cur[1][c] = un[0][c];
                    cur[2][c] = vn[0][c];
#pragma unroll
                    for (int j = 0; j + 1 < K; ++j) {
                        pA[j][c] = pA[j + 1][c];
                        uA[j][c] = uA[j + 1][c];
                        vA[j][c] = vA[j + 1][c];
                        pB[j][c] = pB[j + 1][c];
                        uB[j][c] = uB[j + 1][c];
                        vB[j][c] = vB[j + 1][c];
                        un[j][c] = un[j + 1][c];
                        vn[j][c] = vn[j + 1][c];
                    }
                    pA[K - 1][c] = keep_p;
                    uA[K - 1][c] = keep_u;
                    vA[K - 1][c] = keep_v;
                    pB[K - 1][c] = old_pa;
                    uB[K - 1][c] = old_ua;
                    vB[K - 1][c] = old_va;
                    un[K - 1][c] = nu[c];
                    vn[K - 1][c] = nv[c];
                }
                {
                    const RowTag t0 = info[0], t1 = info[1];
#pragma unroll
                    for (int j = 0; j + 2 < W; ++j) info[j] = info[j + 2];
                    info[W - 2] = t0;
                    info[W - 1] = t1;
                }
            }
            {
                const RowTag first = info[0];
#pragma unroll
                for (int j = 0; j + 1 < W; ++j) info[j] = info[j + 1];
                info[W - 1] = first;
            }

            store_row(cur, r - kLag, cell_r - kLag * nx);   // row r-2K at level K
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
