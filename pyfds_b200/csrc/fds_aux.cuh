// Passes that sit either side of the leapfrog step: the medium-flow row shift of
// AcousticFlow2D and decimated field snapshots for visualisation: pure data movement. (The interactions
// of coupled fields live in fds_couple.cuh.)
#pragma once

#include "fds_common.cuh"

namespace fds {

// ---- AcousticFlow2D.apply_flow (pyfds/acoustic_flow.py:49-57) ---------------------------------------
// After leapfrog step `step`, every grid row n with step % flow_t_deltas[n] == 0 moves one cell towards
// +x:  row[1:nx] = row[0:nx-1]; row[0] = 0  for pressure, velocity_x and velocity_y. `periods` holds
// |flow_t_deltas| per owned row (0 = every step: numpy's `step % 0` is 0). One CTA per (row, component)
// moves its row in place, chunk by chunk from the high end: a chunk is read completely (into registers)
// before any of it is written, and the chunks below it are still untouched when they are read.
constexpr int kFlowThreads = 256;
constexpr int kFlowPerThread = 8;

struct FlowArgs {
    double *state[3];           // origin (local cell 0) of the three components, current buffer
    const long long *periods;   // [rows]
    long long nx, rows;
    long long step;
};

__global__ void __launch_bounds__(kFlowThreads) flow_shift_kernel(FlowArgs a) {
    const long long row = blockIdx.x;
    const long long period = a.periods[row];
    if (period > 1 && a.step % period != 0) return;
    double *__restrict__ line = a.state[blockIdx.y] + row * a.nx;
    constexpr long long kChunk = (long long)kFlowThreads * kFlowPerThread;
    for (long long hi = a.nx; hi > 0; hi -= kChunk) {
        const long long lo = hi > kChunk ? hi - kChunk : 0;
        double v[kFlowPerThread];
#pragma unroll
        for (int j = 0; j < kFlowPerThread; ++j) {
            const long long x = lo + threadIdx.x + (long long)j * kFlowThreads;
            v[j] = (x < hi && x > 0) ? line[x - 1] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kFlowPerThread; ++j) {
            const long long x = lo + threadIdx.x + (long long)j * kFlowThreads;
            if (x < hi) line[x] = v[j];
        }
    }
}

// ---- field snapshots: the frames `Animator._sim_function` puts on its queue (pyfds/gfx.py:72-86) ----
// frame[fy][fx] = values[(fy * stride_y) * nx + fx * stride_x]: every stride-th sample of the owned rows.
struct SnapshotArgs {
    const double *state;   // origin of the component, current buffer
    double *frame;
    long long nx;
    long long fx, fy;      // frame width and height
    int stride_x, stride_y;
};

__global__ void snapshot_kernel(SnapshotArgs a) {
    const long long n = a.fx * a.fy;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long y = k / a.fx, x = k - y * a.fx;
        a.frame[k] = __ldg(a.state + y * a.stride_y * a.nx + x * a.stride_x);
    }
}

}  // namespace fds
