// Interactions between the member fields of a SynchronizedFields group (pyfds/coupling.py:81-300,
// pyfds/coupled_fields.py:45-65), evaluated on the device between two common time steps so that the
// state of the group never leaves the GPU. One thread per cell; every expression is the reference's
// NumPy expression operation for operation (IEEE RN multiply / add / divide, never fused), so the
// linear couplings, the viscous-heating term and the coefficient re-assembly are bit-identical to it.
// The exponential law goes through the device's exp(), which may differ from NumPy's in the last place.
#pragma once

#include "fds_common.cuh"

namespace fds {

constexpr int kCoupleThreads = 256;

// The kernels of a group are launched with programmatic stream serialisation: each grid is scheduled
// while its predecessor still runs and waits for it here (launch latency off the critical path of a
// step that consists of two or three tiny dependent launches).
__device__ __forceinline__ void couple_grid_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// BoundaryCoupling.apply (pyfds/coupling.py:118-140) for a transfer function f given per cell by `Fn`:
//   if accumulate: acc += f(src)                      (acc starts as the integer 0: 0 + x)
//   if step % stepping == 0:
//       delivery = accumulate ? acc : f(src);  target = additive ? target + delivery : delivery
//       acc = 0
struct DeliverArgs {
    double *target;        // component of the target field, current buffer
    double *acc;           // [n] or nullptr (accumulate is False)
    long long n;
    int additive;
    int deliver;           // step % stepping == 0
};

template <typename Fn>
__device__ __forceinline__ void deliver_cell(const DeliverArgs &d, long long k, Fn f) {
    double delivery;
    if (d.acc) {
        const double sum = add(d.acc[k], f(k));
        if (!d.deliver) {
            d.acc[k] = sum;
            return;
        }
        delivery = sum;
        d.acc[k] = 0.0;
    } else {
        if (!d.deliver) return;
        delivery = f(k);
    }
    d.target[k] = d.additive ? add(d.target[k], delivery) : delivery;
}

// transfer function  values -> scale * values
__global__ void couple_linear_kernel(DeliverArgs d, const double *__restrict__ source, double scale) {
    couple_grid_sync();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.n) return;
    deliver_cell(d, k, [&](long long i) { return mul(scale, source[i]); });
}

// ThermoAcoustic1D._loss_coupling (pyfds/coupled_fields.py:45-65):
//   derivative = a_v_p.dot(velocity * density)      a_v_p = backward d_x with factors g = dt/dx/density
//                                                   (row i: (0 + (-g[i-1]) * w[i-1]) + g[i] * w[i])
//   return absorption / density_thermal / heat_capacity * derivative ** 2 * dt
//        = ((gain * (derivative * derivative)) * dt)   with gain baked on the host
struct HeatingArgs {
    const double *velocity;
    const double *density, *g, *gain;   // [n] each
    double dt;
};

__global__ void couple_viscous_heating_kernel(DeliverArgs d, HeatingArgs h) {
    couple_grid_sync();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.n) return;
    deliver_cell(d, k, [&](long long i) {
        const double w = mul(h.velocity[i], h.density[i]);
        // the first row of the band has no sub-diagonal entry: 0 + g*w
        const double below = i > 0 ? mul(-h.g[i - 1], mul(h.velocity[i - 1], h.density[i - 1])) : 0.0;
        const double derivative = add(acc0(below), mul(h.g[i], w));
        return mul(mul(h.gain[i], mul(derivative, derivative)), h.dt);
    });
}

// MaterialCoupling.apply (pyfds/coupling.py:199-215) with one of the built-in laws:
//   exponential (pyfds/coupling.py:250-259):  a + (1 - a) * exp(b * q)
//   power law   (pyfds/coupling.py:291-300):  1 + factor * q ** power
// followed, when the factors moved by more than the threshold (always, without one), by the
// re-assembly of the target field's coefficients with  parameter = static_parameter * factors
// (pyfds/coupling.py:193-197) -- here: per-cell coefficient arrays of the 1-D kernels.
enum { kLawExponential = 0, kLawPower = 1 };
enum { kTargetAcoustic1D = 0, kTargetThermal1D = 1 };

struct LawArgs {
    const double *source;
    double *factors;                 // [n] scratch: factors of this step
    double *last;                    // [n] factors of the last re-assembly (0 before the first)
    unsigned long long *max_bits;    // max |(f - last) / f| as the bits of a non-negative double
    int *reassemblies;               // count, for the host
    long long n;
    int law;
    double p0, p1;                   // exponential: a, b (p2 = 1 - a);  power: factor, power
    double p2;
    int has_threshold;
    double threshold;
    // re-assembly
    int target_model, parameter;     // which of the three parameters is scaled
    const double *statics;           // [3][n] static parameter vectors of the target field
    double *cell_tab;                // [FDS_TAB_COUNT][n] coefficient arrays of the target context
    double k_dtdx, k_dtdx2, k_inv_dx;   // dt / dx, dt / dx ** 2, 1 / dx as the host evaluates them
};

__device__ __forceinline__ double law_factor(const LawArgs &a, double q) {
    if (a.law == kLawExponential) return add(a.p0, mul(a.p2, exp(mul(a.p1, q))));
    double powered;
    // NumPy's fast paths of `array ** scalar` (exact); anything else is pow()
    if (a.p1 == 2.0) powered = mul(q, q);
    else if (a.p1 == 1.0) powered = q;
    else if (a.p1 == 0.5) powered = sqrt(q);
    else if (a.p1 == -1.0) powered = 1.0 / q;
    else powered = pow(q, a.p1);
    return add(1.0, mul(a.p0, powered));
}

__global__ void couple_law_factors_kernel(LawArgs a) {
    couple_grid_sync();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double change = 0.0;
    if (k < a.n) {
        const double f = law_factor(a, a.source[k]);
        a.factors[k] = f;
        change = fabs(sub(f, a.last[k]) / f);
    }
    if (!a.has_threshold) return;
    // max over the block, then one atomic per block (non-negative doubles order like their bits)
    __shared__ unsigned long long block_max[kCoupleThreads / 32];
    unsigned long long bits = (unsigned long long)__double_as_longlong(change);
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if ((threadIdx.x & 31) == 0) block_max[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kCoupleThreads / 32; ++w) bits = block_max[w] > bits ? block_max[w] : bits;
        atomicMax(a.max_bits, bits);
    }
}

__global__ void couple_law_assemble_kernel(LawArgs a) {
    couple_grid_sync();
    if (a.has_threshold) {
        const double change = __longlong_as_double((long long)*a.max_bits);
        if (!(change > a.threshold)) return;
    }
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) atomicAdd(a.reassemblies, 1);
    if (k >= a.n) return;
    const double f = a.factors[k];
    a.last[k] = f;
    double p[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        p[j] = a.statics[(long long)j * a.n + k];
        if (j == a.parameter) p[j] = mul(p[j], f);
    }
    double *tab = a.cell_tab;
    const long long n = a.n;
    if (a.target_model == kTargetAcoustic1D) {
        // pyfds/acoustics.py:27-38:  dt / dx * c ** 2 * rho,  dt / dx / rho,  dt / dx ** 2 * mu / rho
        const double c = p[0], rho = p[1], mu = p[2];
        const double h = mul(a.k_dtdx2, mu) / rho;
        tab[FDS_TAB_FX * n + k] = mul(mul(a.k_dtdx, mul(c, c)), rho);
        tab[FDS_TAB_GX * n + k] = a.k_dtdx / rho;
        tab[FDS_TAB_VM1 * n + k] = h;
        tab[FDS_TAB_V0 * n + k] = mul(-2.0, h);
        tab[FDS_TAB_VP1 * n + k] = h;
    } else {
        // pyfds/thermal.py:33-37:  dt / dx / rho / cp,  1 / dx * kx
        const double rho = p[0], cp = p[1], kx = p[2];
        tab[FDS_TAB_FX * n + k] = (a.k_dtdx / rho) / cp;
        tab[FDS_TAB_GX * n + k] = mul(a.k_inv_dx, kx);
    }
}

}  // namespace fds
