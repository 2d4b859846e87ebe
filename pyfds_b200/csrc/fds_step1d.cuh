// 1-D kernels (Acoustic1D pyfds/acoustics.py:40-52, Thermal1D pyfds/thermal.py:40-51).
//
// A 1-D problem is a few thousand cells stepped tens of thousands of times: the state fits on chip and
// the cost is synchronisation, not bandwidth. Each CTA therefore loads a tile of the line plus a halo
// of `halo` cells per side into shared memory, advances it `n_steps` time steps there (boundaries,
// sources and probes included) and stores the owned cells. Per step the region of valid values
// shrinks by two cells per side (the lossy operator has reach 2), so n_steps <= halo / 2 when the line
// is split over several CTAs; a line that fits one CTA has no neighbours and can take any number of
// steps in a single launch. Cells outside the line are the zero padding with the void material.
#pragma once

#include "fds_common.cuh"

namespace fds {

struct Step1DArgs {
    const double *in[2];   // origin at cell 0
    double *out[2];
    long long n;           // cells of the line
    int tile;              // owned cells per CTA
    int halo;              // extra cells loaded on either side (>= 2)
    int n_steps;           // steps advanced by this launch
    long long sig_index;   // first step - sig_first_step
    long long ring_row;    // probe record of the first step
};

constexpr int k1DThreads = 1024;      // upper bound; the launch uses one thread per cell where it can
constexpr int k1DMaxPerThread = 12;   // (tile + 2*halo) <= 12 * 1024 cells, 18 bytes each

// PER = cells per thread (1 or k1DMaxPerThread). A step is a chain of four short phases separated by
// block barriers, so its cost is instruction issue + latency of one SM, not bandwidth: the plan keeps
// tiles at a few hundred cells with one cell per thread (PER = 1, blockDim = width rounded up to a
// warp), which leaves every phase a few dozen instructions per warp; PER = 12 only serves tiles forced
// larger than a block.
template <bool THERMAL, bool LOSSY, int PER>
__global__ void __launch_bounds__(k1DThreads, 1) step1d_kernel(Step1DArgs a, StepTables t) {
    const int nthreads = blockDim.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int width = a.tile + 2 * a.halo;
    double *s = reinterpret_cast<double *>(smem_raw);     // scalar component
    double *u = s + width;                                // vector component
    map_t *id = reinterpret_cast<map_t *>(u + width);      // material id + flags + classes
    __shared__ double tabs[FDS_TAB_COUNT][kMaxMaterials];

    const int tid = threadIdx.x;
    const long long origin = (long long)blockIdx.x * a.tile - a.halo;  // global cell of local 0

    for (int k = tid; k < FDS_TAB_COUNT * kMaxMaterials; k += nthreads)
        (&tabs[0][0])[k] = t.tab[k];
    for (int l = tid; l < width; l += nthreads) {
        const long long g = origin + l;
        s[l] = a.in[0][g];
        u[l] = THERMAL ? 0.0 : a.in[1][g];
        id[l] = t.map[g];
    }
    __syncthreads();

    double unew[PER];
    for (int q = 0; q < a.n_steps; ++q) {
        const long long sig = a.sig_index + q;
        double *__restrict__ record = t.ring + (a.ring_row + q) * t.n_slots;

        // 1. boundaries and probes of the scalar component
        for (int l = tid; l < width; l += nthreads) {
            const unsigned f = id[l];
            if (f & (kFlagBound | kFlagProbe | kClassMask)) {
                const long long g = origin + l;
                double v = apply_class(t.cls_alpha, t.cls_value, 0, f, s[l]);
                if (f & kFlagBound)
                    v = apply_bounds(t.bound[0], t.rows, t.signals, t.sig_steps, sig, g, v);
                s[l] = v;
                if ((f & kFlagProbe) && l >= a.halo && l < a.halo + a.tile)
                    write_probes(t.probe[0], t.rows, record, g, v);
            }
        }
        __syncthreads();

        // 2. vector component: backward difference of the scalar (+ viscous second difference)
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int l = tid + r * nthreads;
            if (l >= 2 && l < width - 2 && (id[l] & kIdMask)) {
                const int m = id[l] & kIdMask, mm = id[l - 1] & kIdMask;
                const double d = diff2(tabs[FDS_TAB_GX][mm], s[l - 1], tabs[FDS_TAB_GX][m], s[l]);
                if (THERMAL) {
                    unew[r] = -d;
                } else if (LOSSY) {
                    const int mp = id[l + 1] & kIdMask;
                    double vis = acc0(mul(tabs[FDS_TAB_VM1][mm], u[l - 1]));
                    vis = add(vis, mul(tabs[FDS_TAB_V0][m], u[l]));
                    vis = add(vis, mul(tabs[FDS_TAB_VP1][mp], u[l + 1]));
                    unew[r] = sub(u[l], sub(d, vis));
                } else {
                    unew[r] = sub(u[l], d);
                }
            }
        }
        if (LOSSY) __syncthreads();

        // 3. boundaries and probes of the vector component
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int l = tid + r * nthreads;
            if (l >= 2 && l < width - 2 && (id[l] & kIdMask)) {
                const unsigned f = id[l];
                double v = unew[r];
                if (f & (kFlagBound | kFlagProbe | kClassMask)) {
                    const long long g = origin + l;
                    v = apply_class(t.cls_alpha, t.cls_value, 1, f, v);
                    if (f & kFlagBound)
                        v = apply_bounds(t.bound[1], t.rows, t.signals, t.sig_steps, sig, g, v);
                    if ((f & kFlagProbe) && l >= a.halo && l < a.halo + a.tile)
                        write_probes(t.probe[1], t.rows, record, g, v);
                }
                u[l] = v;
            }
        }
        __syncthreads();

        // 4. scalar component: forward difference of the vector component
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int l = tid + r * nthreads;
            if (l >= 2 && l < width - 2 && (id[l] & kIdMask)) {
                const int m = id[l] & kIdMask, mp = id[l + 1] & kIdMask;
                s[l] = sub(s[l], diff2(tabs[FDS_TAB_FX][m], u[l], tabs[FDS_TAB_FX][mp], u[l + 1]));
            }
        }
        // no barrier: phase 1 of the next step only touches a thread's own cells
    }
    __syncthreads();

    for (int l = a.halo + tid; l < a.halo + a.tile; l += nthreads) {
        const long long g = origin + l;
        if (g < a.n) {
            a.out[0][g] = s[l];
            a.out[1][g] = u[l];
        }
    }
}

}  // namespace fds
