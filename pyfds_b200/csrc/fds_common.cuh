// Shared device-side definitions of the B200 finite-difference step engine.
//
// Arithmetic contract: the reference evaluates every operator as a scipy DIA
// mat-vec, i.e. per output cell  y = 0; y += d0*u0; y += d1*u1; ...  in stored-diagonal order, each
// operation one IEEE-754 double RN multiply or add, never fused. All kernels therefore use the
// explicit __dmul_rn/__dadd_rn intrinsics (which nvcc never contracts into DFMA) in exactly that order.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fdsb200.h"

namespace fds {

constexpr int kMaxMaterials = 64;         // 6-bit material id in the per-cell map byte
constexpr uint8_t kIdMask = 0x3F;
constexpr uint8_t kFlagBound = 0x40;      // some component has a boundary operation on this cell
constexpr uint8_t kFlagProbe = 0x80;      // some component has a probe on this cell

// CSR table of boundary operations for one component (fds_upload_boundaries).
struct BoundTable {
    const long long *cells;
    const int *offsets;
    const double *alpha;
    const double *value;
    const int *signal;
    int n_cells;
};

// Probe points of one component (fds_upload_probes).
struct ProbeTable {
    const long long *cells;
    const int *slots;
    int n;
};

// Everything a step kernel needs besides the state pointers. Passed by value (kernel parameter).
struct StepTables {
    const uint8_t *map;        // material id + flags, origin at local cell 0
    const double *tab;         // [FDS_TAB_COUNT][kMaxMaterials]
    const double *ctab;        // [FDS_CTAB_COUNT][n_materials + 1][nx]   (axisymmetric only)
    const double *cvec;        // [FDS_CVEC_COUNT][nx]                    (axisymmetric only)
    BoundTable bound[3];
    ProbeTable probe[3];
    const double *signals;     // [n_signals][sig_steps]
    long long sig_steps;
    long long sig_first_step;
    double *ring;              // probe records [ring_steps][n_slots]
    int n_slots;
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
// "y = 0; y += t": the accumulator of a DIA mat-vec starts at +0.0, so a first product of -0.0 turns
// into +0.0. Kept explicit so that signed zeros match the reference as well.
__device__ __forceinline__ double acc0(double t) { return __dadd_rn(0.0, t); }

// Two-diagonal DIA mat-vec rows  (0 + (-f_a)*u_a) + f_b*u_b   (d_x / d_y, forward or backward;
// pyfds/fields.py:289-297,317-326).
__device__ __forceinline__ double diff2(double fa, double ua, double fb, double ub) {
    return add(acc0(mul(-fa, ua)), mul(fb, ub));
}

// Applies all boundary operations of `cell` in list order:  v = alpha*v + value
// (pyfds/regions.py:136-145 applied by pyfds/fields.py:598-600).
__device__ __noinline__ double apply_bounds(const BoundTable t, const double *__restrict__ signals,
                                            long long sig_steps, long long sig_index,
                                            long long cell, double v) {
    int lo = 0, hi = t.n_cells;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t.cells[mid] < cell) lo = mid + 1; else hi = mid;
    }
    if (lo < t.n_cells && t.cells[lo] == cell) {
        const int end = t.offsets[lo + 1];
        for (int o = t.offsets[lo]; o < end; ++o) {
            const int s = t.signal[o];
            const double val = s >= 0 ? signals[(long long)s * sig_steps + sig_index] : t.value[o];
            v = add(mul(t.alpha[o], v), val);
        }
    }
    return v;
}

// Stores `v` into every probe slot attached to `cell` (pyfds/fields.py:606-611).
__device__ __noinline__ void write_probes(const ProbeTable t, double *__restrict__ record,
                                          long long cell, double v) {
    int lo = 0, hi = t.n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t.cells[mid] < cell) lo = mid + 1; else hi = mid;
    }
    for (; lo < t.n && t.cells[lo] == cell; ++lo) record[t.slots[lo]] = v;
}

}  // namespace fds
