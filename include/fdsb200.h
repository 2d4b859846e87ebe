/*
 * fdsb200.h -- C ABI of the B200 finite-difference step engine (libfdsb200.so).
 *
 * This is the drop-in boundary for the time-stepping hot path of emtpb/pyfds. The reference has no
 * FFI of its own: the seam is the template-method contract of `Field` (pyfds/fields.py:59-95) --
 * `assemble_matrices()` builds scipy DIA operators, `sim_step()` applies them, `simulate()` loops.
 * Each entry point below names the reference code it replaces. Plain pointers and sizes only; all
 * host buffers are borrowed for the duration of the call; device memory belongs to the context.
 *
 * Every function returns 0 on success and a non-zero code on failure; the message is available from
 * fds_last_error(). Nothing here falls back to the CPU: without a CUDA device fds_create() fails.
 *
 * Cell addressing: a context owns `rows` consecutive grid rows starting at global row `row0` of an
 * nx-wide grid (a y-slab; the whole grid on one GPU). A *local cell index* is
 * `x + (y - row0) * nx`; indices in [-halo_rows*nx, 0) and [rows*nx, (rows+halo_rows)*nx) address the
 * neighbour slabs' rows kept as halo. 1-D models use ny = rows = 1.
 */
#ifndef FDSB200_H
#define FDSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fds_ctx fds_ctx;

/* Physical model = which `sim_step` is executed. */
enum fds_model {
    FDS_ACOUSTIC1D = 1,   /* pyfds/acoustics.py:40-52   */
    FDS_ACOUSTIC2D = 2,   /* pyfds/acoustics.py:111-128 */
    FDS_ACOUSTIC3DAXI = 3,/* pyfds/acoustics.py:205-225 */
    FDS_THERMAL1D = 4,    /* pyfds/thermal.py:40-51     */
    FDS_THERMAL2D = 5,    /* pyfds/thermal.py:92-107    */
    FDS_THERMAL3DAXI = 6  /* pyfds/thermal.py:160-176   */
};

/* Field components in device order: the scalar quantity and the two staggered vector components. */
enum fds_component {
    FDS_COMP_S = 0,  /* pressure / temperature            */
    FDS_COMP_X = 1,  /* velocity(_x) / heat_flux(_x)      */
    FDS_COMP_Y = 2   /* velocity_y / heat_flux_y (2-D)    */
};

/* Per-material coefficient tables (index = material id, id 0 = "void": all zero, used for padding).
 * They hold what `assemble_matrices` puts on the operator diagonals (pyfds/acoustics.py:27-38,
 * 89-109,178-203; pyfds/thermal.py:30-38,75-90,139-158), evaluated on the host with the reference's
 * NumPy expressions so that every coefficient has identical bits. */
enum fds_table {
    FDS_TAB_GX = 0,  /* a_vx_p factor  dt/dx/rho           | thermal: Kx = 1/dx*kx (a_qx_t)          */
    FDS_TAB_GY = 1,  /* a_vy_p factor  dt/dy/rho           | thermal: Ky = 1/dy*ky (a_qy_t)          */
    FDS_TAB_FX = 2,  /* a_p_vx factor  dt/dx*c^2*rho       | thermal: Ax = dt/dx/rho/cp (a_t_qx)     */
    FDS_TAB_FY = 3,  /* a_p_vy factor  dt/dy*c^2*rho       | thermal: Ay = dt/dy/rho/cp (a_t_qy)     */
    FDS_TAB_VM1 = 4, /* a_vx_vx diagonal at offset -1  (0 + dt/dx^2*mu/rho)                         */
    FDS_TAB_VP1 = 5, /* a_vx_vx diagonal at offset +1                                                */
    FDS_TAB_VMN = 6, /* a_vx_vx diagonal at offset -nx (0 + dt/dy^2*mu/rho)                         */
    FDS_TAB_VPN = 7, /* a_vx_vx diagonal at offset +nx                                               */
    FDS_TAB_V0 = 8,  /* a_vx_vx main diagonal (0 + -2*Hx) + -2*Hy                                    */
    FDS_TAB_EB = 9,  /* axisymmetric: dt*mu/rho of the extra vx/r^2 term (pyfds/acoustics.py:213-215)*/
    FDS_TAB_COUNT = 10
};

/* Per-(material, column) tables of the axisymmetric models, where 1/r enters the coefficient
 * (pyfds/acoustics.py:181-201, pyfds/thermal.py:141-143). Layout [n_materials + 1][nx]. */
enum fds_column_table {
    FDS_CTAB_FX = 0,  /* a_p_vx / a_t_qx factor divided by r */
    FDS_CTAB_VM1 = 1, /* a_vx_vx offset -1 diagonal incl. the central-difference 1/r term */
    FDS_CTAB_VP1 = 2, /* a_vx_vx offset +1 diagonal incl. the central-difference 1/r term */
    FDS_CTAB_COUNT = 3
};

/* Per-column vectors of the axisymmetric models (`_radii`, pyfds/acoustics.py:166-176). */
enum fds_column_vector {
    FDS_CVEC_R = 0,   /* r  = x + dx/2 */
    FDS_CVEC_RR = 1,  /* r*r           */
    FDS_CVEC_COUNT = 2
};

typedef struct fds_desc {
    int32_t model;        /* enum fds_model */
    int32_t device;       /* CUDA device ordinal */
    int64_t nx;           /* samples along x = row length */
    int64_t ny;           /* global number of rows (1 for 1-D models) */
    int64_t row0;         /* first grid row owned by this context */
    int64_t rows;         /* number of rows owned */
    int32_t halo_rows;    /* rows of neighbour data kept on either side (0 on a single GPU) */
    int32_t lossy;        /* non-zero if any material has absorption_coef != 0 (acoustic models) */
    int32_t n_materials;  /* real materials, ids 1..n_materials (<= 31) */
    int32_t kernel;       /* 0 = automatic, 1 = one-step kernel on global memory, 3 = one-step tile kernel, 2 = force the
                             streaming multi-step kernel (fails if the model/grid does not support it) */
} fds_desc;

/* --- life cycle ------------------------------------------------------------------------------ */

/* Allocates all device state for one slab (zero-initialised). Replaces the implicit state held by
 * `FieldComponent.values` (pyfds/fields.py:585) and the matrices of `assemble_matrices`. */
int fds_create(const fds_desc *desc, fds_ctx **out);
void fds_destroy(fds_ctx *ctx);
/* Message of the last failure on this context (or of the last failed fds_create if ctx is NULL). */
const char *fds_last_error(const fds_ctx *ctx);
/* Number of CUDA devices visible (0 if the driver is missing); never fails. */
int fds_device_count(void);

/* --- setup: what `assemble_matrices()` produces (pyfds/acoustics.py:89-109 etc.) ------------- */

/* Material id of every local cell, halo rows included: n = (rows + 2*halo_rows) * nx entries starting
 * at local cell -halo_rows*nx. Replaces `Field.material_vector` painting (pyfds/fields.py:34-57). */
int fds_upload_material_map(fds_ctx *ctx, const uint8_t *ids, int64_t n);
/* One per-material table: n = n_materials + 1 doubles (entry 0 is the void material). */
int fds_upload_table(fds_ctx *ctx, int32_t table, const double *values, int64_t n);
/* One per-(material, column) table: n = (n_materials + 1) * nx doubles. */
int fds_upload_column_table(fds_ctx *ctx, int32_t table, const double *values, int64_t n);
/* One per-column vector: n = nx doubles. */
int fds_upload_column_vector(fds_ctx *ctx, int32_t vec, const double *values, int64_t n);
/* Fields whose material parameters differ from cell to cell (what `MaterialCoupling` produces by
 * scaling a parameter with another field's values, pyfds/coupling.py:182-197): one per-CELL coefficient
 * array, n = owned cells, used instead of the per-material table of the same id. All tables the model
 * reads must be uploaded this way; values == NULL switches back to the per-material tables. 1-D models
 * keep their kernel; plain 2-D models (single slab, not axisymmetric) run on the one-thread-per-cell
 * kernel, one step per launch, while per-cell coefficients are on. */
int fds_upload_cell_table(fds_ctx *ctx, int32_t table, const double *values, int64_t n);

/* --- boundaries and sources: `FieldComponent.apply_bounds` + `Boundary.apply`
 *     (pyfds/fields.py:591-600, pyfds/regions.py:125-145) -------------------------------------- */

/* Boundary operations of one component in CSR form: `cells` are the n_cells distinct local cell
 * indices (ascending) that carry at least one operation; the operations of cells[k] are
 * offsets[k] .. offsets[k+1]-1, stored in the order of the component's `boundaries` list. Operation o
 * computes  v = alpha[o] * v + (signal[o] < 0 ? value[o] : signals[signal[o]][step]).
 * Cells in halo rows must be included (they are recomputed redundantly). */
int fds_upload_boundaries(fds_ctx *ctx, int32_t component, const int64_t *cells,
                          const int32_t *offsets, int64_t n_cells, const double *alpha,
                          const double *value, const int32_t *signal, int64_t n_ops);
/* Signal window: row-major [n_signals][n_steps] samples for absolute steps first_step ..
 * first_step + n_steps - 1 (`value[step]`, pyfds/regions.py:141,144). */
int fds_upload_signals(fds_ctx *ctx, const double *samples, int64_t n_signals, int64_t n_steps,
                       int64_t first_step);

/* --- probes: `FieldComponent.write_outputs` (pyfds/fields.py:602-611) ------------------------ */

/* Probe points of one component: cells ascending (owned cells only; duplicates allowed), slots[k] is
 * the column of the probe record that receives cells[k]. n_slots_total is the record width shared by
 * all components (pass the same value for each component). */
int fds_upload_probes(fds_ctx *ctx, int32_t component, const int64_t *cells, const int32_t *slots,
                      int64_t n, int64_t n_slots_total);

/* --- state: `FieldComponent.values` (pyfds/fields.py:585) ----------------------------------- */

/* Copies n = rows * nx owned values of one component from / to host memory. */
int fds_upload_state(fds_ctx *ctx, int32_t component, const double *values, int64_t n);
int fds_download_state(fds_ctx *ctx, int32_t component, double *values, int64_t n);
/* Optional: page-locks a host array that will be used repeatedly with fds_upload_state /
 * fds_download_state (`values` arrays are updated in place by the reference, pyfds/acoustics.py:117),
 * so that the copies run as direct DMA. The caller must unregister before the memory is freed. */
int fds_host_register(void *host, int64_t bytes);
int fds_host_unregister(void *host);
/* Zeroes all components on the device (`Field.reset`, pyfds/fields.py:121-127). */
int fds_reset_state(fds_ctx *ctx);

/* --- the hot path: `Field.simulate` loop body (pyfds/fields.py:87-93) = `sim_step` x n_steps -- */

/* Advances the slab by n_steps steps starting at absolute step first_step. probes_out (may be NULL if
 * there are no probes) receives row-major [n_steps][n_slots_total] samples; slots not owned by this
 * slab are left untouched. The signal window uploaded before must cover the step range.
 * Synchronous: returns after the state and all probe records are complete. */
int fds_step(fds_ctx *ctx, int64_t first_step, int64_t n_steps, double *probes_out);
/* Same, but only enqueues the work on the context's stream (no probe drain to the host, probes_out
 * semantics do not apply); pair with fds_sync(). Used for device-resident timing. */
int fds_step_async(fds_ctx *ctx, int64_t first_step, int64_t n_steps);
int fds_sync(fds_ctx *ctx);

/* `Field.simulate(n)` of a field that lives in host arrays (pyfds/fields.py:67-95) in ONE call:
 * values_in[c] -> device, n_steps steps, device -> values_out[c] (c < 2 for 1-D, 3 for 2-D models; owned
 * cells each; in and out may be the same arrays), probe records to probes_out as fds_step does. For
 * short calls on large 2-D grids the three phases are overlapped by row bands: band j is uploaded while
 * the steps advance the rows uploaded so far as far as their dependency cone allows and finished rows
 * are downloaded -- the call then takes about as long as the slower PCIe direction instead of the sum of
 * both. Results are bit-identical to upload + fds_step + download. Page-lock the arrays
 * (fds_host_register) for the copies to run asynchronously. Single-slab contexts. */
int fds_simulate(fds_ctx *ctx, int64_t first_step, int64_t n_steps, const double *const *values_in,
                 double *const *values_out, double *probes_out);
/* Row bands the last fds_simulate call was cut into (0: it ran the three phases back to back). */
int fds_last_pipeline_bands(fds_ctx *ctx, int64_t *bands);

/* --- the step after the hot path: medium flow ------------------------------- */

/* `AcousticFlow2D.apply_flow` (pyfds/acoustic_flow.py:49-57) on the device: after leapfrog step s every
 * owned row n with s % periods[n] == 0 is moved one cell towards +x (row[1:] = row[:-1], row[0] = 0) in
 * all three components. periods[n] = |flow_t_deltas[row0 + n]| (pyfds/acoustic_flow.py:34); 0 and 1
 * both mean "after every step" (numpy evaluates s % 0 to 0). n = ny: the periods of ALL grid rows --
 * every slab of a multi-GPU run passes the same array, so that all slabs end their launches at the same
 * steps (a single-slab context may pass its n = rows periods). The step kernels keep advancing several
 * steps per launch and end a launch where a row of the grid has to move. periods == NULL or n == 0
 * switches the flow off. Acoustic2D only. */
int fds_set_flow(fds_ctx *ctx, const int64_t *periods, int64_t n);
/* Number of row-shift passes the last fds_step / fds_step_async call launched. */
int fds_last_flow_shifts(fds_ctx *ctx, int64_t *shifts);

/* --- the caller above the hot path: coupled fields --------------------------- */

/* `SynchronizedFields.sim_step` (pyfds/coupling.py:81-87) for 1-D member fields on one device: per
 * common step every member advances one step (in list order), then the interactions run (in the order
 * they were added) -- all as kernels on one stream, so the state of the group stays in HBM for the
 * whole call. Interactions are the reference's own ones whose transfer function is built in or
 * linear; arbitrary Python transfer functions stay with the host-side session of the Python layer. */
typedef struct fds_group fds_group;
int fds_group_create(fds_ctx *const *members, int32_t n_members, fds_group **out);
void fds_group_destroy(fds_group *group);
const char *fds_group_last_error(const fds_group *group);
/* `BoundaryCoupling` (pyfds/coupling.py:90-140) with transfer function  values -> scale * values:
 * target (+)= scale * source every `stepping`-th step; with `accumulate` the products are summed after
 * every step and the sum is delivered. accumulated: the coupling's `accumulated_transfer` carried over
 * from an earlier call (n doubles) or NULL for 0. */
int fds_group_add_linear(fds_group *group, int32_t src_member, int32_t src_component,
                         int32_t dst_member, int32_t dst_component, double scale, int32_t additive,
                         int32_t accumulate, int64_t stepping, const double *accumulated);
/* `ThermoAcoustic1D._loss_coupling` (pyfds/coupled_fields.py:45-65): temperature +=
 * gain * (a_v_p . (velocity * density))^2 * dt, with per-cell density, gradient_factor (the factors
 * dt/dx/density of a_v_p, pyfds/acoustics.py:31) and gain = absorption_coef / density_thermal /
 * heat_capacity, n doubles each, evaluated by the host with the reference's expressions. */
int fds_group_add_viscous_heating(fds_group *group, int32_t sound_member, int32_t heat_member,
                                  const double *density, const double *gradient_factor,
                                  const double *gain, double dt, int32_t accumulate,
                                  int64_t stepping, const double *accumulated);
/* `MaterialCouplingExponential` (law 0: a = p0, b = p1) / `MaterialCouplingPowerLaw` (law 1:
 * factor = p0, power = p1), pyfds/coupling.py:218-300: every `stepping`-th step the factors of the
 * source component are evaluated and -- if they moved by more than `threshold` relative to the ones
 * last used, or always without a threshold -- the target member's per-cell coefficients are
 * re-assembled from  statics[parameter] * factors  (pyfds/coupling.py:193-215). statics: the three
 * static parameter vectors of the target in the order the model lists them ([3][n] doubles: acoustic
 * sound_velocity, density, absorption_coef; thermal density, heat_capacity, thermal_conductivity_x);
 * last: `last_used_factors` of an earlier call or NULL; scalars: dt/dx, dt/dx**2, 1/dx as the host
 * evaluates them. The target needs per-cell coefficient tables (fds_upload_cell_table). */
int fds_group_add_material_law(fds_group *group, int32_t src_member, int32_t src_component,
                               int32_t dst_member, int32_t parameter, int32_t law, double p0,
                               double p1, int32_t has_threshold, double threshold,
                               int64_t stepping, const double *statics, const double *last,
                               const double *scalars);
/* n_steps common steps; probes_out[m] receives member m's records [n_steps][n_slots] (may be NULL for
 * members without probes). Synchronous. */
int fds_group_step(fds_group *group, int64_t first_step, int64_t n_steps, double *const *probes_out);
/* State an interaction carries between calls: the accumulated transfer of a boundary coupling, or the
 * factors last used by a material law (n doubles) and how often it re-assembled so far. */
int fds_group_read(fds_group *group, int32_t interaction, double *values, int64_t *reassemblies);

/* --- the output side: field snapshots --------------------------------------- */

/* What `Animator._sim_function` puts on its queue after every `steps_per_frame` steps
 * (pyfds/gfx.py:72-86: `getattr(field, observed_component).values`), without moving the whole state
 * to the host: every stride_x-th sample of every stride_y-th owned row of one component, taken from
 * the state as it is after the steps enqueued so far, is gathered into frame slot 0 or 1 on the device
 * and copied to page-locked host memory on a second stream while the caller's next steps run.
 * Frame size: ceil(nx / stride_x) * ceil(rows / stride_y) doubles, row-major. */
int fds_snapshot_async(fds_ctx *ctx, int32_t component, int32_t stride_x, int32_t stride_y,
                       int32_t slot);
/* Waits for the snapshot in `slot` and copies its n doubles to `frame`. */
int fds_snapshot_wait(fds_ctx *ctx, int32_t slot, double *frame, int64_t n);

/* --- multi-GPU: y-slab halo exchange over NCCL (no reference equivalent) ------- */

/* Writes a 128-byte NCCL unique id (rank 0 calls this, the host side broadcasts it). */
int fds_comm_unique_id(uint8_t id[128]);
/* Joins the communicator; slab `rank` exchanges halo rows with rank-1 and rank+1. */
int fds_comm_init(fds_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world);

/* Optional peer-memory halo path: after every launch the outermost rows of a slab are copied straight
 * into the neighbour slabs' halo rows over NVLink (their buffers are mapped through CUDA IPC) and the
 * launches of adjacent slabs are ordered by flags in peer memory -- no ncclSend/ncclRecv, no separate
 * band launches, no host involvement in the time loop. fds_peer_export writes
 * 7 CUDA IPC handles of 64 bytes (six state buffers, one flag block); every rank passes the handles of
 * rank-1 (side 0) and rank+1 (side 1) to fds_peer_import together with that neighbour's row count. */
int fds_peer_export(fds_ctx *ctx, uint8_t *handles);
int fds_peer_import(fds_ctx *ctx, int32_t side, const uint8_t *handles, int64_t neighbour_rows);

/* The same for slabs that live in ONE process (one context per GPU, e.g. one host thread each): no NCCL
 * communicator and no IPC -- fds_slab_init names the slab's place in the row partition, fds_peer_connect
 * hands it the neighbour context itself (peer access is enabled between the two devices). At the start
 * of every step call a slab copies the neighbours' edge rows into its halo rows (all slabs must have
 * finished their uploads before any of them steps); after that the rows travel from inside the step
 * kernels as above. The slabs' step calls must run concurrently (they wait for each other on the
 * device). */
/* Everything a following fds_step over the same steps would allocate or load (flag tables, probe ring
 * and pinned buffer, strip census, task tables, the kernels themselves), without stepping. Slabs of one
 * process call it before the first of them starts stepping: with peer access enabled a device
 * allocation synchronises with the peer devices, and a neighbour that is already waiting inside a step
 * kernel for this slab would never let it return. */
int fds_step_prepare(fds_ctx *ctx, int64_t first_step, int64_t n_steps);
int fds_slab_init(fds_ctx *ctx, int32_t rank, int32_t world);
int fds_peer_connect(fds_ctx *ctx, int32_t side, fds_ctx *neighbour);

/* --- measurement ------------------------------------------------------------------------------ */

/* Device time in milliseconds of the step kernels launched by the last fds_step/fds_step_async call,
 * measured with CUDA events on the launching stream; valid after fds_sync(). */
int fds_last_step_ms(fds_ctx *ctx, double *ms);
/* Number of step-kernel launches issued by the last fds_step/fds_step_async call and their name. */
int fds_last_launch_info(fds_ctx *ctx, int64_t *launches, int64_t *steps_per_launch,
                         const char **kernel_name);
/* Bytes of device memory held by the context. */
int64_t fds_device_bytes(const fds_ctx *ctx);
/* Diagnostics of the streaming kernel, accumulated since fds_create when the environment variable
 * FDS_STREAM_STATS is set (all zero otherwise): out[0..4] entries into the branch-free row-pair body
 * by variant (no boundary operation, constant operations on component 0 / 1 / 2, several materials),
 * out[5] rows that took the general row iteration, out[6] rows streamed in total, out[7] row pairs of
 * the lossy axisymmetric model that were stepped with the IEEE division because a numerator was
 * outside the range of the quotient sequence (fds_streamv.cuh). Test support: no reference
 * counterpart. */
int fds_stream_stats(fds_ctx *ctx, int64_t out[8]);

#ifdef __cplusplus
}
#endif

#endif /* FDSB200_H */
