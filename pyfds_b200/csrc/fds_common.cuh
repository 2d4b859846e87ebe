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

// Per-cell map entry (16 bit):
//   bits 0-4   material id (0 = void, 1..31)
//   bit  5     some component has boundary operations that need the table lookup (slow path)
//   bit  6     some component has a probe on this cell
//   bits 7-9   class of the constant boundary operation on component 0 (0 = none)
//   bits 10-12 ... on component 1,  bits 13-15 ... on component 2
// A cell whose boundary list for a component is exactly ONE scalar operation (a Dirichlet wall, a
// constant additive source) gets a class: v = alpha[class]*v + value[class] is applied inline, without
// any table access. Everything else (signals, several operations on one cell) takes the slow path.
typedef uint16_t map_t;
constexpr int kMaxMaterials = 32;
constexpr unsigned kIdMask = 0x1F;
constexpr unsigned kFlagBound = 0x20;
constexpr unsigned kFlagProbe = 0x40;
constexpr unsigned kClassMask = 0xFF80;
constexpr int kClassBits = 3;
constexpr int kMaxClasses = 8;            // class 0 = none
__host__ __device__ constexpr int class_shift(int comp) { return 7 + kClassBits * comp; }

// CSR table of boundary operations for one component (fds_upload_boundaries). `row_ptr[k]` is the
// first entry of `cells` that lies in local row k (counted from the first halo row), so a lookup only
// searches the handful of entries of one grid row.
struct BoundTable {
    const long long *cells;
    const int *offsets;
    const double *alpha;
    const double *value;
    const int *signal;
    const int *row_ptr;
    int n_cells;
};

// Probe points of one component (fds_upload_probes), with the same per-row index.
struct ProbeTable {
    const long long *cells;
    const int *slots;
    const int *row_ptr;
    int n;
};

// 1-D fields: what the boundary and probe tables hold for one flagged cell and component, resolved on
// the host when the tables change, so that a launch reads it with two loads instead of walking the
// tables (binary searches = a chain of dependent global loads per flagged cell and launch, which is what
// a one-step launch of a coupled group used to spend most of its time on).
struct LineResolved {
    double alpha, value;   // first operation: v = alpha * v + (signal sample or value)
    int signal;            // signal of the first operation, -1 = scalar value
    int o0, n_ops;         // all operations: o0 .. o0 + n_ops - 1 of the component's table
    int p0, p1;            // probe entries p0 .. p1 - 1 of the component's table
    int pad;
};

// Geometry needed to find the row of a cell on the slow path.
struct RowIndex {
    long long nx;
    long long halo_cells;   // cells between the first halo row and local cell 0
};

// Everything a step kernel needs besides the state pointers. Passed by value (kernel parameter).
struct StepTables {
    const map_t *map;          // material id + flags + classes, origin at local cell 0
    const double *tab;         // [FDS_TAB_COUNT][kMaxMaterials]
    const double *ctab;        // [FDS_CTAB_COUNT][n_materials + 1][nx]   (axisymmetric only)
    const double *cvec;        // [FDS_CVEC_COUNT][nx]                    (axisymmetric only)
    // 1-D fields whose materials vary from cell to cell (MaterialCoupling, pyfds/coupling.py:143-215):
    // per-cell coefficients [FDS_TAB_COUNT][cell_n] instead of the per-material table (null otherwise)
    const double *cell_tab;
    long long cell_n;
    // 1-D: entry number of every cell (-1: none) and the entries [number][component]
    const int *line_index;
    const LineResolved *line_entries;
    BoundTable bound[3];
    ProbeTable probe[3];
    const double *signals;     // [n_signals][sig_steps]
    long long sig_steps;
    long long sig_first_step;
    double *ring;              // probe records [ring_steps][n_slots]
    int n_slots;
    RowIndex rows;
    double cls_alpha[3][kMaxClasses];   // single boundary operations by class: v = alpha * v + value,
    double cls_value[3][kMaxClasses];   // value constant, or (2-D models) sample `step` of signal
    int cls_signal[3][kMaxClasses];     // cls_signal (-1: constant)
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

// Inline constant boundary operation of component `comp` (class bits of the map entry).
__device__ __forceinline__ double apply_class(const double (*alpha)[kMaxClasses],
                                              const double (*value)[kMaxClasses], int comp,
                                              unsigned entry, double v) {
    const unsigned k = (entry >> class_shift(comp)) & (kMaxClasses - 1);
    return k ? add(mul(alpha[comp][k], v), value[comp][k]) : v;
}

// Value of class k of component comp at signal sample `sig` (index into the uploaded window).
__device__ __forceinline__ double class_value(const StepTables &t, int comp, unsigned k,
                                              long long sig) {
    const int sidx = t.cls_signal[comp][k];
    return sidx >= 0 ? __ldg(t.signals + (long long)sidx * t.sig_steps + sig) : t.cls_value[comp][k];
}

// Position of `cell` in the sorted array `cells[lo, hi)` or the position of the first larger entry.
// Boundary lines and rectangles are runs of consecutive cells, so the entry is usually exactly
// `cell - cells[lo]` places after `lo`; otherwise fall back to a binary search.
__device__ __forceinline__ int lower_bound_cell(const long long *__restrict__ cells, int lo, int hi,
                                                long long cell) {
    if (lo >= hi) return hi;
    const long long guess = lo + (cell - __ldg(cells + lo));
    if (guess >= lo && guess < hi && __ldg(cells + guess) == cell) {
        // duplicates (probes) sit next to each other: step back to the first one
        int g = (int)guess;
        while (g > lo && __ldg(cells + g - 1) == cell) --g;
        return g;
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cells + mid) < cell) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Applies all boundary operations of `cell` in list order:  v = alpha*v + value
// (pyfds/regions.py:136-145 applied by pyfds/fields.py:598-600).
__device__ __noinline__ double apply_bounds(const BoundTable t, const RowIndex ri,
                                            const double *__restrict__ signals, long long sig_steps,
                                            long long sig_index, long long cell, double v) {
    if (t.n_cells == 0) return v;
    const long long row = (cell + ri.halo_cells) / ri.nx;
    const int hi = __ldg(t.row_ptr + row + 1);
    const int k = lower_bound_cell(t.cells, __ldg(t.row_ptr + row), hi, cell);
    if (k < hi && __ldg(t.cells + k) == cell) {
        const int end = __ldg(t.offsets + k + 1);
        for (int o = __ldg(t.offsets + k); o < end; ++o) {
            const int s = __ldg(t.signal + o);
            const double val =
                s >= 0 ? __ldg(signals + (long long)s * sig_steps + sig_index) : __ldg(t.value + o);
            v = add(mul(__ldg(t.alpha + o), v), val);
        }
    }
    return v;
}

// Stores `v` into every probe slot attached to `cell` (pyfds/fields.py:606-611).
__device__ __noinline__ void write_probes(const ProbeTable t, const RowIndex ri,
                                          double *__restrict__ record, long long cell, double v) {
    if (t.n == 0) return;
    const long long row = (cell + ri.halo_cells) / ri.nx;
    const int hi = __ldg(t.row_ptr + row + 1);
    int k = lower_bound_cell(t.cells, __ldg(t.row_ptr + row), hi, cell);
    for (; k < hi && __ldg(t.cells + k) == cell; ++k) record[__ldg(t.slots + k)] = v;
}

// ---- multi-GPU halo rows over NVLink peer memory ----------------------------------------------------
// After every launch the outermost `halo` rows of a slab are copied straight into the halo rows of the
// neighbour slabs (their buffers are mapped through CUDA IPC) and a flag in the neighbour's memory is
// released with the launch number; before the next launch a one-thread kernel waits for the flags the
// neighbours wrote. No host involvement, no NCCL call in the time loop.
struct HaloPushArgs {
    const double *src[6];    // up to 2 sides x 3 components
    double *dst[6];
    long long count;         // doubles per segment (multiple of 2)
    int n_segments;
    unsigned *flag_out[2];   // neighbours' flags (peer memory), nullptr if there is no neighbour
    unsigned launch_id;
    unsigned *blocks_done;   // zero before the launch; reset by the last block
};

__global__ void halo_push_kernel(HaloPushArgs a) {
    const int seg = blockIdx.y;
    const double2 *__restrict__ src = reinterpret_cast<const double2 *>(a.src[seg]);
    double2 *__restrict__ dst = reinterpret_cast<double2 *>(a.dst[seg]);
    const long long n2 = a.count / 2;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n2;
         k += (long long)gridDim.x * blockDim.x)
        dst[k] = src[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(a.blocks_done, 1u) == total - 1u) {
            *a.blocks_done = 0u;
            __threadfence_system();
            for (int side = 0; side < 2; ++side)
                if (a.flag_out[side])
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag_out[side]),
                                 "r"(a.launch_id)
                                 : "memory");
        }
    }
}

__global__ void halo_wait_kernel(const unsigned *flag_lo, const unsigned *flag_hi, unsigned want) {
    const unsigned *flags[2] = {flag_lo, flag_hi};
    for (int side = 0; side < 2; ++side) {
        if (!flags[side]) continue;
        unsigned seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];"
                         : "=r"(seen)
                         : "l"(flags[side])
                         : "memory");
            if ((int)(seen - want) < 0) __nanosleep(100);
        } while ((int)(seen - want) < 0);
    }
}

}  // namespace fds
