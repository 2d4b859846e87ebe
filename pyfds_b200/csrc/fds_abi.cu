// C ABI of the B200 finite-difference step engine (see include/fdsb200.h for the contract and the
// reference code every entry point replaces).
//
// HBM layout of one context (one y-slab):
//   * every field component lives in TWO buffers (steps are out of place and ping-pong between them);
//     each buffer is  [ pad | halo rows | owned rows | halo rows | pad ]  doubles, zero-initialised.
//     The pad (>= 9 rows + 4096 cells) is never written: cells outside the global grid read as 0.0 with
//     the void material, which makes every skipped DIA term an exact "+ 0.0" and removes all edge
//     special-casing from the kernels (and gives the row-wrap of the x operators for free, because
//     cells are addressed by flat index).
//   * one byte per cell: material id (6 bits) | boundary flag | probe flag, same padding.
//   * coefficient tables: [FDS_TAB_COUNT][64] doubles (+ per-column tables for axisymmetric models).
//   * boundary operations / probe points as small sorted tables, signals as a [n_signals][n_steps]
//     window, probe records in a two-half device ring drained to pinned host memory on a second
//     stream while the next half is being computed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler injects its library

#include "fds_common.cuh"
#include "fds_step1d.cuh"
#include "fds_line1d.cuh"
#include "fds_step2d.cuh"
#include "fds_stream2d.cuh"
#include "fds_streamv.cuh"
#include "fds_aux.cuh"
#include "fds_couple.cuh"

using namespace fds;

namespace {

// NVTX range for the duration of a scope (bake / upload / step / drain show up as
// named ranges on an nsys or ncu --nvtx timeline).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

std::mutex g_err_mutex;
std::string g_create_error = "";

struct DevArray {
    void *ptr = nullptr;
    size_t bytes = 0;
};

// ---- NCCL, resolved lazily so that single-GPU use has no dependency on it ----------------------
struct ncclComm;
typedef struct { char internal[128]; } nccl_unique_id;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(ncclComm **, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(ncclComm *) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8;  // ncclDouble

const char *load_nccl() {
    if (g_nccl.lib) return nullptr;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return "libnccl.so.2 not found";
#define FDS_NCCL_SYM(field, name)                                                        \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                                 \
    if (!g_nccl.field) { g_nccl.lib = nullptr; return "NCCL symbol " name " missing"; }
    FDS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    FDS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    FDS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FDS_NCCL_SYM(Send, "ncclSend")
    FDS_NCCL_SYM(Recv, "ncclRecv")
    FDS_NCCL_SYM(GroupStart, "ncclGroupStart")
    FDS_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FDS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FDS_NCCL_SYM
    return nullptr;
}

}  // namespace

struct fds_ctx {
    fds_desc d{};
    int dims = 2;
    int ncomp = 3;
    bool thermal = false;
    bool axi = false;
    long long owned = 0;      // owned cells
    long long halo = 0;       // halo cells per side
    long long pad = 0;        // pad cells per side (beyond the halo)
    long long alloc = 0;      // cells per buffer
    double *buf[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    map_t *map = nullptr;
    bool map_uploaded = false;
    double *tab = nullptr;
    double *cell_tab = nullptr;   // 1-D: per-cell coefficients [FDS_TAB_COUNT][nx] (fds_upload_cell_table)
    double *ctab = nullptr;
    double *cvec = nullptr;
    int cur = 0;

    DevArray bcells[3], boffsets[3], balpha[3], bvalue[3], bsignal[3], browptr[3];
    long long n_bcells[3] = {0, 0, 0};
    DevArray pcells[3], pslots[3], prowptr[3];
    long long n_probes[3] = {0, 0, 0};
    std::vector<long long> host_bcells[3];   // host copies: strip weights of the streaming kernel
    std::vector<long long> host_ccells[3];
    // cells whose single scalar operation is applied inline: per component the cells and their class
    DevArray ccells[3], cclass[3];
    long long n_ccells[3] = {0, 0, 0};
    double cls_alpha[3][kMaxClasses] = {}, cls_value[3][kMaxClasses] = {};
    int cls_signal[3][kMaxClasses];   // signal behind a class's value, -1 = constant (set by fds_create)
    // task tables of the streaming kernels, one per (row range, steps per launch) seen so far, and
    // the census they are balanced with: rows per strip that cannot take the branch-free body
    struct StreamPlan {
        long long row_begin = 0, row_end = 0;
        int k = 0, n_tasks = 0;
        void *tasks = nullptr;
        // ordering of consecutive sweeps over this table (TaskSync in fds_stream2d.cuh)
        int *dep_ptr = nullptr, *dep_idx = nullptr;
        unsigned *done = nullptr;
        int n_edge[2] = {0, 0};   // tasks that own rows of the lower / upper edge band
    };
    std::vector<StreamPlan> plans;
    // Task tables live in an arena (device memory + page-locked staging of the same size): building one
    // costs neither a cudaMalloc nor a synchronisation, which matters when fds_simulate builds dozens
    // of them between launches while copies are in flight.
    struct PlanChunk {
        char *dev = nullptr, *host = nullptr;
        size_t capacity = 0, used = 0;
    };
    std::vector<PlanChunk> plan_chunks;
    std::vector<int> strip_nonplain;   // [strip][block of kCensusBlockRows rows]
    int census_blocks = 0;
    bool census_valid = false;
    DevArray flagged;         // cells whose flag bits are currently set
    long long n_flagged = 0;
    bool flags_dirty = false;
    DevArray signals;
    long long sig_steps = 0, sig_first = 0, n_signals = 0;

    long long n_slots = 0;
    DevArray ring;
    long long ring_half = 0;  // steps per ring half
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    std::vector<int> host_pslots[3];   // host copies of the probe slot lists
    // 1-D: host copies of the lookup tables and what refresh_flags resolves from them (LineResolved)
    std::vector<int> host_boffsets[3], host_bsignal[3];
    std::vector<double> host_balpha[3], host_bvalue[3];
    std::vector<long long> host_pcells[3];
    DevArray line_index, line_entries;

    cudaStream_t stream = nullptr, drain = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_half[2] = {nullptr, nullptr}, ev_drained[2] = {nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_edge = nullptr, ev_comm = nullptr;
    bool timed = false;

    long long last_launches = 0, last_steps_per_launch = 0;
    const char *last_kernel = "none";
    long long device_bytes = 0;

    ncclComm *comm = nullptr;
    int rank = 0, world = 1;

    // peer-memory halo path of the streaming kernel (multi-GPU): neighbours' buffers opened over IPC
    unsigned *flags = nullptr;           // [0],[1]: halo-arrival flags written by the lower / upper
                                         // neighbour; [2]: blocks of halo_push_kernel that are done;
                                         // [4..7]: finished edge tasks by side and sweep parity;
                                         // [8]: a device-side wait gave up
    void *peer_base[2][2][3] = {};       // [side][buffer][component] base pointers (IPC mappings)
    unsigned *peer_flags[2] = {nullptr, nullptr};
    long long peer_rows[2] = {0, 0};
    bool peer_open[2] = {false, false};
    // slabs of ONE process (fds_peer_connect): the neighbour contexts themselves; no IPC, no NCCL
    fds_ctx *local_neighbour[2] = {nullptr, nullptr};
    bool local_peers = false;
    bool dry_run = false;     // fds_step_prepare: everything a launch needs, but no launch
    unsigned launch_seq = 0;
    // task table of the streaming launch that was enqueued last on `stream`, nothing else since
    // (nullptr otherwise): the next launch over the same table may overlap it
    const void *chain_tasks = nullptr;
    bool overlap = true;       // programmatic dependent launches between sweeps (FDS_NO_OVERLAP=1: off)
    bool halo_in_kernel = true;   // halo rows pushed / awaited inside the streaming kernels
    bool kernels_before_cells[3] = {false, false, false};   // use_stream2d / use_streamv / use_tile2d
                                                            // while 2-D per-cell coefficients are on
    bool use_stream2d = false; // streaming multi-step kernel selected
    bool use_tile2d = false;   // shared-memory tile kernel selected (one step per launch)
    bool use_streamv = false;  // streaming kernel of the viscous / axisymmetric acoustic models
    StepTables *d_tables = nullptr;   // device copy of the tables for the streaming kernel's slow path
    unsigned long long *stream_stats = nullptr;   // FDS_STREAM_STATS: counters of the streaming kernel
    int sv_ctas = 0;   // resident CTAs per SM of the viscous kernel: 0 = by model (launch_streamv),
                       // FDS_SV_CTAS = 2 or 3 overrides
    int *task_counters = nullptr;     // pool of zeroed work counters, one per streaming launch
    int next_counter = 0;
    int chunk_rows = 0;       // rows per streaming task (0 = heuristic)
    int tile_rows = 0;        // owned rows per tile of the tile kernel (0 = default)
    int halo_1d = 0;          // halo cells of the tiled 1-D kernel (0 = default)
    int max_k = kMaxStreamSteps;

    // medium flow (AcousticFlow2D): |flow_t_deltas| of every owned row and their distinct values
    DevArray flow_periods;
    std::vector<long long> flow_unique;
    bool flow = false;
    long long flow_shifts = 0;        // flow_shift_kernel launches of the last call

    // field snapshots: two frame slots, each a device buffer and a pinned host buffer
    double *frame_dev[2] = {nullptr, nullptr};
    size_t frame_dev_cap[2] = {0, 0};
    void *frame_host[2] = {nullptr, nullptr};
    size_t frame_host_cap[2] = {0, 0};
    long long frame_n[2] = {0, 0};
    cudaEvent_t ev_frame_ready[2] = {nullptr, nullptr}, ev_frame_done[2] = {nullptr, nullptr};

    // fds_simulate: per row band one event "uploaded" and one "final rows stored"
    std::vector<cudaEvent_t> band_events;
    long long last_bands = 0;          // bands the last fds_simulate call was cut into (0: not pipelined)

    std::string err;
};

namespace {

int fail(fds_ctx *ctx, const std::string &msg) {
    if (ctx) ctx->err = msg;
    else {
        std::lock_guard<std::mutex> lock(g_err_mutex);
        g_create_error = msg;
    }
    return 1;
}

#define FDS_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

#define FDS_NCCL(ctx, call)                                                              \
    do {                                                                                 \
        int e_ = (call);                                                                 \
        if (e_ != 0)                                                                     \
            return fail(ctx, std::string(#call) + ": " + g_nccl.GetErrorString(e_));     \
    } while (0)

int dev_alloc(fds_ctx *ctx, void **p, size_t bytes, bool zero) {
    FDS_CUDA(ctx, cudaMalloc(p, std::max<size_t>(bytes, 16)));
    if (zero) FDS_CUDA(ctx, cudaMemsetAsync(*p, 0, std::max<size_t>(bytes, 16), ctx->stream));
    ctx->device_bytes += (long long)bytes;
    return 0;
}

int dev_upload(fds_ctx *ctx, DevArray &a, const void *host, size_t bytes) {
    if (a.bytes < bytes || !a.ptr) {
        if (a.ptr) {
            FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(a.ptr);
            ctx->device_bytes -= (long long)a.bytes;
        }
        a.ptr = nullptr;
        a.bytes = 0;
        void *p = nullptr;
        if (dev_alloc(ctx, &p, bytes, false)) return 1;
        a.ptr = p;
        a.bytes = bytes;
    }
    if (bytes)
        FDS_CUDA(ctx, cudaMemcpyAsync(a.ptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    // the host buffer is borrowed only for the duration of the call
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Sets / clears bits of map entries; `values` (optional) holds a per-cell value that is shifted into
// place (boundary-operation classes). Entries are 16 bit, the update is atomic on the 32-bit word.
__global__ void flag_kernel(map_t *map, const long long *cells, const int *values, int shift,
                            long long n, unsigned set_mask, unsigned clear_mask) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    map_t *entry = map + cells[k];
    unsigned int *word =
        reinterpret_cast<unsigned int *>(reinterpret_cast<uintptr_t>(entry) & ~uintptr_t(3));
    const int pos = 8 * (int)(reinterpret_cast<uintptr_t>(entry) & 3);
    unsigned set = set_mask;
    if (values) set |= (unsigned)values[k] << shift;
    if (clear_mask) atomicAnd(word, ~(clear_mask << pos));
    if (set) atomicOr(word, set << pos);
}

int launch_flags(fds_ctx *ctx, const long long *cells, const int *values, int shift, long long n,
                 unsigned set_mask, unsigned clear_mask) {
    if (n <= 0) return 0;
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    flag_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        ctx->map + ctx->pad + ctx->halo, cells, values, shift, n, set_mask, clear_mask);
    FDS_CUDA(ctx, cudaGetLastError());
    return 0;
}

__global__ void widen_ids_kernel(map_t *map, const uint8_t *ids, long long n) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) map[k] = ids[k];
}

// Per-row index of a sorted cell list: entry k = first position whose cell lies in local row >= k
// (rows counted from the first halo row).
int upload_row_ptr(fds_ctx *ctx, DevArray &dst, const long long *cells, int64_t n) {
    const long long total_rows = (ctx->dims == 1) ? 1 : ctx->d.rows + 2 * (long long)ctx->d.halo_rows;
    std::vector<int> ptr((size_t)total_rows + 1, 0);
    int64_t k = 0;
    for (long long r = 0; r <= total_rows; ++r) {
        const long long first_cell = r * ctx->d.nx - ctx->halo;   // first cell of row r
        while (k < n && cells[k] < first_cell) ++k;
        ptr[(size_t)r] = (int)k;
    }
    ptr[(size_t)total_rows] = (int)n;
    return dev_upload(ctx, dst, ptr.data(), ptr.size() * sizeof(int));
}

// Rebuilds the per-cell flag bits from the boundary and probe tables.
void invalidate_plans(fds_ctx *ctx);

// 1-D: per flagged cell and component the position of its operations and probes in the tables
// (LineResolved), so that the kernels do not search them launch after launch.
int resolve_line_entries(fds_ctx *ctx) {
    const long long n = ctx->d.nx;
    std::vector<int> index((size_t)n, -1);
    std::vector<LineResolved> entries;
    auto entry_of = [&](long long cell) -> LineResolved * {
        int &number = index[(size_t)cell];
        if (number < 0) {
            number = (int)(entries.size() / 2);
            LineResolved blank{1.0, 0.0, -1, 0, 0, 0, 0, 0};
            entries.push_back(blank);
            entries.push_back(blank);
        }
        return &entries[(size_t)(2 * number)];
    };
    for (int c = 0; c < 2; ++c) {
        const auto &cells = ctx->host_bcells[c];
        for (size_t k = 0; k < cells.size(); ++k) {
            if (cells[k] < 0 || cells[k] >= n) continue;
            LineResolved &e = entry_of(cells[k])[c];
            e.o0 = ctx->host_boffsets[c][k];
            e.n_ops = ctx->host_boffsets[c][k + 1] - e.o0;
            if (e.n_ops > 0) {
                e.alpha = ctx->host_balpha[c][(size_t)e.o0];
                e.value = ctx->host_bvalue[c][(size_t)e.o0];
                e.signal = ctx->host_bsignal[c][(size_t)e.o0];
            }
        }
        const auto &probes = ctx->host_pcells[c];
        for (size_t k = 0; k < probes.size();) {
            size_t end = k;
            while (end < probes.size() && probes[end] == probes[k]) ++end;
            if (probes[k] >= 0 && probes[k] < n) {
                LineResolved &e = entry_of(probes[k])[c];
                e.p0 = (int)k;
                e.p1 = (int)end;
            }
            k = end;
        }
    }
    if (dev_upload(ctx, ctx->line_index, index.data(), index.size() * sizeof(int))) return 1;
    if (dev_upload(ctx, ctx->line_entries, entries.data(), entries.size() * sizeof(LineResolved)))
        return 1;
    return 0;
}

int refresh_flags(fds_ctx *ctx) {
    if (!ctx->flags_dirty) return 0;
    invalidate_plans(ctx);   // the strip census reads the flag and class bits
    const unsigned all_bits = kFlagBound | kFlagProbe | kClassMask;
    if (ctx->n_flagged)
        if (launch_flags(ctx, (const long long *)ctx->flagged.ptr, nullptr, 0, ctx->n_flagged, 0,
                         all_bits))
            return 1;
    long long total = 0;
    for (int c = 0; c < 3; ++c) total += ctx->n_bcells[c] + ctx->n_probes[c] + ctx->n_ccells[c];
    if (ctx->flagged.bytes < (size_t)total * 8) {
        if (ctx->flagged.ptr) {
            FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->flagged.ptr);
            ctx->device_bytes -= (long long)ctx->flagged.bytes;
            ctx->flagged = DevArray();
        }
        void *p = nullptr;
        if (dev_alloc(ctx, &p, (size_t)total * 8, false)) return 1;
        ctx->flagged.ptr = p;
        ctx->flagged.bytes = (size_t)total * 8;
    }
    long long at = 0;
    long long *dst = (long long *)ctx->flagged.ptr;
    for (int c = 0; c < 3; ++c) {
        const long long nb = ctx->n_bcells[c], np = ctx->n_probes[c], nc = ctx->n_ccells[c];
        if (nb) {
            FDS_CUDA(ctx, cudaMemcpyAsync(dst + at, ctx->bcells[c].ptr, nb * 8,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
            if (launch_flags(ctx, dst + at, nullptr, 0, nb, kFlagBound, 0)) return 1;
            at += nb;
        }
        if (np) {
            FDS_CUDA(ctx, cudaMemcpyAsync(dst + at, ctx->pcells[c].ptr, np * 8,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
            if (launch_flags(ctx, dst + at, nullptr, 0, np, kFlagProbe, 0)) return 1;
            at += np;
        }
        if (nc) {
            FDS_CUDA(ctx, cudaMemcpyAsync(dst + at, ctx->ccells[c].ptr, nc * 8,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
            if (launch_flags(ctx, dst + at, (const int *)ctx->cclass[c].ptr, class_shift(c), nc, 0, 0))
                return 1;
            at += nc;
        }
    }
    ctx->n_flagged = total;
    ctx->flags_dirty = false;
    if (ctx->dims == 1 && resolve_line_entries(ctx)) return 1;
    return 0;
}


// ---- host <-> device state copies -----------------------------------------------------------------
// One DMA per array. NumPy arrays are pageable; the host side may page-lock the arrays it reuses
// (fds_host_register) so that these copies run at PCIe rate instead of through the driver's bounce
// buffers.
int state_copy(fds_ctx *ctx, void *device, void *host, size_t bytes, bool to_device) {
    if (to_device)
        FDS_CUDA(ctx, cudaMemcpyAsync(device, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    else
        FDS_CUDA(ctx, cudaMemcpyAsync(host, device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

double *origin(fds_ctx *ctx, int which, int comp) {
    return ctx->buf[which][comp] + ctx->pad + ctx->halo;
}

StepTables make_tables(fds_ctx *ctx) {
    StepTables t{};
    t.map = ctx->map + ctx->pad + ctx->halo;
    t.tab = ctx->tab;
    t.ctab = ctx->ctab;
    t.cvec = ctx->cvec;
    t.cell_tab = ctx->cell_tab;
    t.cell_n = ctx->owned;
    t.line_index = (const int *)ctx->line_index.ptr;
    t.line_entries = (const LineResolved *)ctx->line_entries.ptr;
    for (int c = 0; c < 3; ++c) {
        t.bound[c].cells = (const long long *)ctx->bcells[c].ptr;
        t.bound[c].offsets = (const int *)ctx->boffsets[c].ptr;
        t.bound[c].alpha = (const double *)ctx->balpha[c].ptr;
        t.bound[c].value = (const double *)ctx->bvalue[c].ptr;
        t.bound[c].signal = (const int *)ctx->bsignal[c].ptr;
        t.bound[c].row_ptr = (const int *)ctx->browptr[c].ptr;
        t.bound[c].n_cells = (int)ctx->n_bcells[c];
        t.probe[c].cells = (const long long *)ctx->pcells[c].ptr;
        t.probe[c].slots = (const int *)ctx->pslots[c].ptr;
        t.probe[c].row_ptr = (const int *)ctx->prowptr[c].ptr;
        t.probe[c].n = (int)ctx->n_probes[c];
    }
    t.signals = (const double *)ctx->signals.ptr;
    t.sig_steps = ctx->sig_steps;
    t.sig_first_step = ctx->sig_first;
    t.ring = (double *)ctx->ring.ptr;
    t.n_slots = (int)ctx->n_slots;
    t.rows.nx = ctx->d.nx;
    t.rows.halo_cells = ctx->halo;
    memcpy(t.cls_alpha, ctx->cls_alpha, sizeof(t.cls_alpha));
    memcpy(t.cls_value, ctx->cls_value, sizeof(t.cls_value));
    memcpy(t.cls_signal, ctx->cls_signal, sizeof(t.cls_signal));
    return t;
}

// ---- kernel dispatch ----------------------------------------------------------------------------

constexpr int kCounterPool = 4096;
constexpr int kFlagWords = 16;      // fds_ctx::flags
constexpr int kFlagEdgeDone = 4, kFlagError = 8;

int *next_counter(fds_ctx *ctx) {
    if (ctx->next_counter == kCounterPool) {
        if (cudaMemsetAsync(ctx->task_counters, 0, sizeof(int) * kCounterPool, ctx->stream) !=
            cudaSuccess)
            return nullptr;
        ctx->chain_tasks = nullptr;   // (the counters of a sweep that may still run must not be cleared)
        ctx->next_counter = 0;
    }
    return ctx->task_counters + ctx->next_counter++;
}

template <int MODEL, bool LOSSY>
int launch_step2d(fds_ctx *ctx, const Step2DArgs &a, const StepTables &t) {
    const long long rows = a.row_end - a.row_begin;
    if (rows <= 0) return 0;
    if (ctx->use_tile2d) {
        constexpr bool thermal = (MODEL == FDS_THERMAL2D || MODEL == FDS_THERMAL3DAXI);
        Tile2DArgs g{};
        g.rows_below = 1;
        g.rows_above = (LOSSY && !thermal) ? 2 : 1;
        g.tile_h = thermal ? 64 : 24;
        if (ctx->tile_rows > 0) g.tile_h = ctx->tile_rows;
        g.tiles_x = (int)((a.nx + kTileW - 1) / kTileW);
        g.n_tiles = (int)(g.tiles_x * ((rows + g.tile_h - 1) / g.tile_h));
        g.counter = next_counter(ctx);
        if (!g.counter) return fail(ctx, "tile2d: counter reset failed");
        const int tile_rows = g.tile_h + g.rows_below + g.rows_above;
        const int smem = tile_rows * kTilePitch * ((thermal ? 1 : 3) * 8 + (int)sizeof(map_t));
        auto kernel = tile2d_kernel<MODEL, LOSSY>;
        static int configured_by_device[64] = {};     // function attributes are per device
        int &configured = configured_by_device[ctx->d.device & 63];
        if (configured < smem) {
            FDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               smem));
            configured = smem;
        }
        const int ctas = std::min(g.n_tiles, 148 * 2);
        if (ctx->dry_run) {
            cudaFuncAttributes loaded;
            FDS_CUDA(ctx, cudaFuncGetAttributes(&loaded, kernel));
            return 0;
        }
        kernel<<<ctas, kTileThreads, smem, ctx->stream>>>(a, t, ctx->d.n_materials + 1, g);
        FDS_CUDA(ctx, cudaGetLastError());
        return 0;
    }
    dim3 grid((unsigned)rows, (unsigned)((a.nx + 255) / 256));
    if (ctx->dry_run) {
        cudaFuncAttributes loaded;
        if (ctx->cell_tab) return 0;
        FDS_CUDA(ctx, cudaFuncGetAttributes(&loaded, step2d_kernel<MODEL, LOSSY>));
        return 0;
    }
    if (ctx->cell_tab) {
        constexpr bool axi = (MODEL == FDS_ACOUSTIC3DAXI || MODEL == FDS_THERMAL3DAXI);
        if (axi) return fail(ctx, "per-cell coefficients are not supported for axisymmetric models");
        step2d_kernel<axi ? FDS_ACOUSTIC2D : MODEL, LOSSY, true>
            <<<grid, 256, 0, ctx->stream>>>(a, t, ctx->d.n_materials + 1);
    } else {
        step2d_kernel<MODEL, LOSSY><<<grid, 256, 0, ctx->stream>>>(a, t, ctx->d.n_materials + 1);
    }
    FDS_CUDA(ctx, cudaGetLastError());
    return 0;
}

int dispatch_step2d(fds_ctx *ctx, const Step2DArgs &a, const StepTables &t) {
    const bool lossy = ctx->d.lossy != 0;
    switch (ctx->d.model) {
        case FDS_ACOUSTIC2D:
            return lossy ? launch_step2d<FDS_ACOUSTIC2D, true>(ctx, a, t)
                         : launch_step2d<FDS_ACOUSTIC2D, false>(ctx, a, t);
        case FDS_ACOUSTIC3DAXI:
            return lossy ? launch_step2d<FDS_ACOUSTIC3DAXI, true>(ctx, a, t)
                         : launch_step2d<FDS_ACOUSTIC3DAXI, false>(ctx, a, t);
        case FDS_THERMAL2D:
            return launch_step2d<FDS_THERMAL2D, false>(ctx, a, t);
        case FDS_THERMAL3DAXI:
            return launch_step2d<FDS_THERMAL3DAXI, false>(ctx, a, t);
    }
    return fail(ctx, "model has no 2-D kernel");
}

const char *step2d_name(const fds_ctx *ctx) {
    if (ctx->use_tile2d) {
        switch (ctx->d.model) {
            case FDS_ACOUSTIC2D: return ctx->d.lossy ? "tile2d_kernel<acoustic2d,lossy>"
                                                     : "tile2d_kernel<acoustic2d,lossless>";
            case FDS_ACOUSTIC3DAXI: return ctx->d.lossy ? "tile2d_kernel<acoustic3daxi,lossy>"
                                                        : "tile2d_kernel<acoustic3daxi,lossless>";
            case FDS_THERMAL2D: return "tile2d_kernel<thermal2d>";
            case FDS_THERMAL3DAXI: return "tile2d_kernel<thermal3daxi>";
        }
    }
    switch (ctx->d.model) {
        case FDS_ACOUSTIC2D: return ctx->d.lossy ? "step2d_kernel<acoustic2d,lossy>"
                                                 : "step2d_kernel<acoustic2d,lossless>";
        case FDS_ACOUSTIC3DAXI: return ctx->d.lossy ? "step2d_kernel<acoustic3daxi,lossy>"
                                                    : "step2d_kernel<acoustic3daxi,lossless>";
        case FDS_THERMAL2D: return "step2d_kernel<thermal2d>";
        case FDS_THERMAL3DAXI: return "step2d_kernel<thermal3daxi>";
    }
    return "none";
}


// ---- streaming multi-step kernel (fds_stream2d.cuh) ---------------------------------------------

bool stream_supported(const fds_desc &d) {
    const bool model_ok = ((d.model == FDS_ACOUSTIC2D || d.model == FDS_ACOUSTIC3DAXI) && !d.lossy) ||
                          d.model == FDS_THERMAL2D || d.model == FDS_THERMAL3DAXI;
    return model_ok && d.nx % 4 == 0 && d.nx >= 128;
}

bool streamv_supported(const fds_desc &d) {
    const bool model_ok = (d.model == FDS_ACOUSTIC2D || d.model == FDS_ACOUSTIC3DAXI) && d.lossy;
    return model_ok && d.nx % 4 == 0 && d.nx >= 128;
}

// steps per launch the streaming kernels support for this model
int stream_max_steps(const fds_ctx *ctx) {
    // the viscous / axisymmetric kernel: the 4-cell strip halo covers two steps of its stencil
    if (ctx->use_streamv) return kMaxStreamVSteps;
    return kMaxStreamSteps;
}

// Rows of every strip that cannot take the branch-free body of the streaming kernels: flagged for
// a table lookup or a probe, constant operations on more than one component (or on any, where the
// row is not of one material), or map words that differ from the row before (the criterion of RowTag::steady and of the pair
// vote in fds_stream2d.cuh): one warp per (strip, row).
constexpr int kCensusBlockRows = 32;   // the census counts per strip and block of rows

// ONE_MATERIAL: rows of several materials never count as steady (fds_streamv.cuh)
template <int C, bool ONE_MATERIAL>   // C = cells per lane
__global__ void strip_census_kernel(const map_t *map, long long nx, long long rows, int n_strips,
                                    int stride, int halo, int n_blocks, int *nonplain) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long item = warp; item < rows * n_strips; item += n_warps) {
        const int strip = (int)(item % n_strips);
        const long long row = item / n_strips;
        const long long x0 = (long long)strip * stride - halo + C * lane;
        unsigned long long raw = 0, prev = 0, unit = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            raw |= (unsigned long long)map[row * nx + x0 + c] << (16 * c);
            prev |= (unsigned long long)map[(row - 1) * nx + x0 + c] << (16 * c);
            unit |= 1ull << (16 * c);
        }
        const unsigned long long first = __shfl_sync(0xffffffffu, raw, 0) & kIdMask;
        const bool relevant = x0 < nx + halo;
        if (!relevant) raw = prev = first * unit;
        unsigned comps = 0;   // components with constant operations in this row
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (__any_sync(0xffffffffu, (raw & ((7ull * unit) << class_shift(c))) != 0))
                comps |= 1u << c;
        const bool uniform = __all_sync(0xffffffffu, (raw & (kIdMask * unit)) == first * unit);
        const bool ok = (raw & ((kFlagBound | kFlagProbe) * unit)) == 0 && (row == 0 || raw == prev);
        const bool classes_ok =
            uniform ? (comps & (comps - 1u)) == 0u : (!ONE_MATERIAL && comps == 0u);
        if ((!__all_sync(0xffffffffu, ok) || !classes_ok) && lane == 0)
            atomicAdd(nonplain + (long long)strip * n_blocks + row / kCensusBlockRows, 1);
    }
}

int strip_census(fds_ctx *ctx, int n_strips) {
    NvtxRange nvtx_range("fds:plan census");
    const int n_blocks = (int)((ctx->d.rows + kCensusBlockRows - 1) / kCensusBlockRows);
    const size_t n = (size_t)n_strips * (size_t)n_blocks;
    int *d_counts = nullptr;
    if (dev_alloc(ctx, (void **)&d_counts, sizeof(int) * n, true)) return 1;
    const map_t *map = ctx->map + ctx->pad + ctx->halo;
    if (ctx->use_streamv && ctx->axi)   // its branch-free body takes rows of one material only
        strip_census_kernel<kS2LaneCells, true><<<148 * 4, 256, 0, ctx->stream>>>(
            map, ctx->d.nx, ctx->d.rows, n_strips, kS2StripStride, kS2StripHalo, n_blocks, d_counts);
    else
        strip_census_kernel<kS2LaneCells, false><<<148 * 4, 256, 0, ctx->stream>>>(
            map, ctx->d.nx, ctx->d.rows, n_strips, kS2StripStride, kS2StripHalo, n_blocks, d_counts);
    FDS_CUDA(ctx, cudaGetLastError());
    ctx->strip_nonplain.assign(n, 0);
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->strip_nonplain.data(), d_counts, sizeof(int) * n,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_counts);
    ctx->device_bytes -= (long long)(sizeof(int) * n);
    ctx->census_blocks = n_blocks;
    ctx->census_valid = true;
    return 0;
}

void invalidate_plans(fds_ctx *ctx) {
    if (!ctx->plans.empty()) {
        cudaStreamSynchronize(ctx->stream);
        ctx->plans.clear();
        for (auto &chunk : ctx->plan_chunks) chunk.used = 0;   // the memory is kept for the next tables
    }
    ctx->chain_tasks = nullptr;
    ctx->census_valid = false;
}

// Resident CTAs per SM of the viscous streaming kernel (see launch_streamv for the measurements).
static int streamv_ctas_per_sm(const fds_ctx *ctx) {
    if (ctx->sv_ctas) return ctx->sv_ctas;
    return ctx->d.model == FDS_ACOUSTIC3DAXI ? 2 : kSVCtasPerSm;
}

// Task table of one streaming launch: every strip is cut into chunks of rows so that all tasks cost
// about the same and fill a whole number of rounds over the 148 SMs x CTAs x warps.
//   cost of a task = its rows + overhead, a row that takes the general row iteration counting
//   general_weight times (census: rows that are not steady, each of which sends about k + 3 rows
//   through the general iteration); overhead = rows streamed in addition to the owned ones
//   (pipeline fill, ring fill).
// The number of rounds R follows from a target task height; the smallest cost C is then found for
// which the tasks fit R * slots. Strips without general rows are cut evenly, the others where the
// accumulated cost crosses the multiples of their share. Tasks of the latter are handed out first,
// the rest chunk-major (dynamic distribution absorbs what the model misses).
int build_stream_plan(fds_ctx *ctx, fds_ctx::StreamPlan &plan, int n_strips, int k, int lag_rows) {
    NvtxRange nvtx_range("fds:plan tasks");
    const long long rows = plan.row_end - plan.row_begin;
    // (the viscous kernel: 12 warps per SM also where it runs with 2 resident CTAs -- config 3 with
    // 2 CTAs: tasks cut for 8 warps per SM 117.8, for 12 warps per SM 121.4 Gcell-updates/s)
    const double slots = 148.0 * (ctx->use_streamv ? kSVCtasPerSm : kS2CtasPerSm) * kStreamWarps;
    const double overhead = 2.0 * lag_rows + 4.0;
    // measured on B200 (4096^2, bench.py, 3 CTAs/SM): 1.5 and 2.5 -> 346, 4 -> 341 Gcell-updates/s.
    double general_weight = 2.5;
    if (const char *env = getenv("FDS_GENERAL_WEIGHT")) general_weight = atof(env);
    const int nb = ctx->census_blocks;
    auto row_cost = [&](int s, long long row) {
        const int count = ctx->strip_nonplain[(size_t)s * nb + (size_t)(row / kCensusBlockRows)];
        const double general =
            std::min<double>(kCensusBlockRows, (double)count * (k + 3)) / kCensusBlockRows;
        return 1.0 + (general_weight - 1.0) * general;
    };
    std::vector<double> strip_cost((size_t)n_strips, (double)rows);
    std::vector<char> has_general((size_t)n_strips, 0);
    for (int s = 0; s < n_strips; ++s) {
        for (long long b = plan.row_begin / kCensusBlockRows;
             b <= (plan.row_end - 1) / kCensusBlockRows; ++b)
            if (ctx->strip_nonplain[(size_t)s * nb + (size_t)b] && general_weight != 1.0)
                has_general[(size_t)s] = 1;
        if (!has_general[(size_t)s]) continue;
        double total = 0;
        for (long long row = plan.row_begin; row < plan.row_end; ++row) total += row_cost(s, row);
        strip_cost[(size_t)s] = total;
    }
    const long long min_rows = std::min<long long>(rows, 8);
    std::vector<long long> count((size_t)n_strips, 1), best_count;
    auto tasks_for = [&](double cost) {
        double n = 0;
        const double room = std::max(cost - overhead, 1.0);
        for (int s = 0; s < n_strips; ++s) {
            long long c = (long long)std::ceil(strip_cost[(size_t)s] / room);
            c = std::max<long long>(1, std::min(c, std::max<long long>(rows / min_rows, 1)));
            count[(size_t)s] = c;
            n += (double)c;
        }
        return n;
    };
    if (ctx->chunk_rows > 0) {
        const long long h = std::min<long long>(ctx->chunk_rows, rows);
        best_count.assign((size_t)n_strips, (rows + h - 1) / h);
    } else {
        // Rounds: SMs do not run at one speed (measured 16384^2 on B200: 1 round of 2048-row tasks 299,
        // 4 rounds 361, 16 rounds of 128-row tasks 401 Gcell-updates/s -- a quarter of the time was
        // the tail of the slowest SMs), so tasks stay about kTargetRows tall and the dynamic
        // distribution evens the SMs out; shorter tasks pay more for the rows streamed twice.
        double total_cost = 0;
        for (int s = 0; s < n_strips; ++s) total_cost += strip_cost[(size_t)s];
        // With two or three rounds the tail of the slowest SMs weighs more than the rows streamed
        // twice: shorter tasks there. Measured (profiles/r1_target_rows.log, 128 / 96 / 64 rows):
        // axisymmetric 8192x4096 (2.7 rounds at 128) 286 / 302 / 304 Gcell-updates/s; lossless 4096^2
        // (1.3 rounds: one task per warp either way) 342.9 / 340.2 in an alternating 2000-step A/B;
        // the viscous kernel (twice the pipeline fill per task) 176 / 163 / 154.
        double target_rows = 128.0;
        const double rounds_at_128 = total_cost / slots / 128.0;
        if (!ctx->use_streamv && rounds_at_128 >= 2.0 && rounds_at_128 < 4.0) target_rows = 96.0;
        if (const char *env = getenv("FDS_TARGET_ROWS")) target_rows = std::max(8.0, atof(env));
        const int rounds = (int)std::max(1.0, std::floor(total_cost / slots / target_rows + 0.5));
        double lo = overhead + (double)min_rows, hi = general_weight * (double)rows + overhead;
        if (tasks_for(lo) <= rounds * slots) hi = lo;
        for (int it = 0; it < 50 && hi - lo > 0.01; ++it) {
            const double mid = 0.5 * (lo + hi);
            if (tasks_for(mid) <= rounds * slots) hi = mid; else lo = mid;
        }
        tasks_for(hi);
        best_count = count;
    }
    struct Item { int strip; long long ys, ye; double cost; };
    std::vector<Item> items;
    // strips with general rows: cut where the accumulated cost crosses the multiples of the share
    for (int s = 0; s < n_strips; ++s) {
        if (!has_general[(size_t)s] || ctx->chunk_rows > 0) continue;
        const long long n = best_count[(size_t)s];
        const double share = strip_cost[(size_t)s] / (double)n;
        double acc = 0, chunk_cost = 0;
        long long ys = plan.row_begin, cuts = 1;
        for (long long row = plan.row_begin; row < plan.row_end; ++row) {
            const double c = row_cost(s, row);
            acc += c;
            chunk_cost += c;
            if (row + 1 == plan.row_end || (acc >= share * (double)cuts && cuts < n)) {
                items.push_back({s, ys, row + 1, chunk_cost + overhead});
                ys = row + 1;
                chunk_cost = 0;
                ++cuts;
            }
        }
    }
    std::stable_sort(items.begin(), items.end(),
                     [](const Item &x, const Item &y) { return x.cost > y.cost; });
    // the other strips: even chunks in chunk-major order -- tasks that are handed out together (the
    // warps of a CTA, neighbouring CTAs) stream neighbouring strips of the same rows, i.e. adjacent
    // memory: DRAM pages and the halo columns shared by two strips are reused while they are hot
    long long max_count = 0;
    for (int s = 0; s < n_strips; ++s) max_count = std::max(max_count, best_count[(size_t)s]);
    // (strip-major, FDS_TASK_ORDER=s, measured 3-4 % slower on every configuration)
    bool strip_major = false;
    if (const char *env = getenv("FDS_TASK_ORDER")) strip_major = env[0] == 's';
    const long long outer = strip_major ? n_strips : max_count, inner = strip_major ? max_count : n_strips;
    for (long long o = 0; o < outer; ++o) {
        for (long long i = 0; i < inner; ++i) {
            const int s = (int)(strip_major ? o : i);
            const long long c = strip_major ? i : o;
            const long long n = best_count[(size_t)s];
            if (c >= n) continue;
            if (ctx->chunk_rows > 0) {   // forced height (tests): fixed chunks, a short last one
                const long long h = std::min<long long>(ctx->chunk_rows, rows);
                const long long ys = plan.row_begin + c * h;
                items.push_back({s, ys, std::min(ys + h, plan.row_end), 0.0});
                continue;
            }
            if (has_general[(size_t)s]) continue;
            const long long base = rows / n, extra = rows % n;   // even split into n chunks
            const long long ys = plan.row_begin + c * base + std::min(c, extra);
            const long long h = base + (c < extra ? 1 : 0);
            items.push_back({s, ys, ys + h, (double)h + overhead});
        }
    }
    // Multi-GPU, whole slab: tasks that read halo rows wait for the neighbour's flag inside the kernel,
    // tasks that own rows of an edge band push them to the neighbour when they finish (TaskSync). The
    // edge tasks go first: their rows travel while the interior is computed, and by the time the
    // neighbour's next sweep asks for them they have long arrived.
    const bool whole_slab = plan.row_begin == 0 && plan.row_end == ctx->d.rows;
    const bool has_side[2] = {ctx->world > 1 && ctx->rank > 0, ctx->world > 1 && ctx->rank < ctx->world - 1};
    const long long band = ctx->d.halo_rows;
    auto flags_of = [&](const Item &it) {
        int f = 0;
        if (!whole_slab) return f;
        if (has_side[0]) {
            if (it.ys - lag_rows < 0) f |= kTaskWaitLo;
            if (it.ys < band) f |= kTaskPushLo;
        }
        if (has_side[1]) {
            if (it.ye + lag_rows > rows) f |= kTaskWaitHi;
            if (it.ye > rows - band) f |= kTaskPushHi;
        }
        return f;
    };
    if (whole_slab && (has_side[0] || has_side[1])) {
        // Short edge tasks: the rows a neighbour waits for are cut off the first / last task of every
        // strip as tasks of their own (twice the band: a few dozen rows of work including the
        // pipeline fill), so that they are stored, pushed and flagged within the first tenth of a
        // sweep instead of at its end -- a neighbour may then lag or lead by most of a sweep before
        // anybody waits.
        const long long edge = std::min<long long>(std::max<long long>(2 * band, 8), rows);
        std::vector<Item> cut;
        cut.reserve(items.size() + 2 * (size_t)n_strips);
        for (const Item &it : items) {
            long long ys = it.ys, ye = it.ye;
            if (has_side[0] && ys < edge && ye > edge + 4) {
                cut.push_back({it.strip, ys, edge, (double)(edge - ys) + overhead});
                ys = edge;
            }
            if (has_side[1] && ye > rows - edge && ys < rows - edge - 4) {
                cut.push_back({it.strip, rows - edge, ye, (double)(ye - (rows - edge)) + overhead});
                ye = rows - edge;
            }
            cut.push_back({it.strip, ys, ye, it.cost});
        }
        items.swap(cut);
        std::stable_partition(items.begin(), items.end(),
                              [&](const Item &it) { return flags_of(it) != 0; });
    }
    std::vector<int4> tasks(items.size());
    plan.n_edge[0] = plan.n_edge[1] = 0;
    for (size_t i = 0; i < items.size(); ++i) {
        const int f = flags_of(items[i]);
        if (f & kTaskPushLo) ++plan.n_edge[0];
        if (f & kTaskPushHi) ++plan.n_edge[1];
        tasks[i] = make_int4(items[i].strip, (int)items[i].ys, (int)items[i].ye, f);
    }
    // Dependency lists for overlapped sweeps: the tasks (of the same table, one sweep earlier) whose
    // output a task reads -- strips s-1, s, s+1 (cyclic: the flat index wraps from one row into the
    // next) and rows within `lag_rows` (+1 for that wrap) of its own. The same tasks are the ones that
    // read what it overwrites, so one wait covers both hazards.
    std::vector<int> dep_ptr(items.size() + 1, 0), dep_idx;
    {
        std::vector<std::vector<int>> by_strip((size_t)n_strips);
        for (size_t i = 0; i < items.size(); ++i) by_strip[(size_t)items[i].strip].push_back((int)i);
        const long long margin = lag_rows + 1;
        for (size_t i = 0; i < items.size(); ++i) {
            const Item &t = items[i];
            int neighbours[3] = {(t.strip + n_strips - 1) % n_strips, t.strip, (t.strip + 1) % n_strips};
            const int n_nb = n_strips >= 3 ? 3 : n_strips;   // fewer than 3 strips: no duplicates
            for (int j = 0; j < n_nb; ++j) {
                const int sj = n_strips >= 3 ? neighbours[j] : j;
                for (int other : by_strip[(size_t)sj]) {
                    const Item &o = items[(size_t)other];
                    if (o.ys < t.ye + margin && o.ye > t.ys - margin) dep_idx.push_back(other);
                }
            }
            dep_ptr[i + 1] = (int)dep_idx.size();
        }
    }
    // one block in the arena: tasks | dep_ptr | dep_idx | done (zero), staged in page-locked memory and
    // sent with one asynchronous copy ahead of the launches that read it (same stream)
    auto align16 = [](size_t b) { return (b + 15) / 16 * 16; };
    const size_t off_ptr = align16(tasks.size() * sizeof(int4));
    const size_t off_idx = off_ptr + align16(dep_ptr.size() * sizeof(int));
    const size_t off_done = off_idx + align16(dep_idx.size() * sizeof(int));
    const size_t bytes = off_done + align16(tasks.size() * sizeof(unsigned));
    fds_ctx::PlanChunk *chunk = nullptr;
    for (auto &c : ctx->plan_chunks)
        if (c.capacity - c.used >= bytes) { chunk = &c; break; }
    if (!chunk) {
        fds_ctx::PlanChunk fresh;
        fresh.capacity = std::max<size_t>(bytes, 8u << 20);
        FDS_CUDA(ctx, cudaMalloc((void **)&fresh.dev, fresh.capacity));
        if (cudaHostAlloc((void **)&fresh.host, fresh.capacity, cudaHostAllocPortable) != cudaSuccess) {
            cudaFree(fresh.dev);
            return fail(ctx, "build_stream_plan: cudaHostAlloc failed");
        }
        ctx->device_bytes += (long long)fresh.capacity;
        ctx->plan_chunks.push_back(fresh);
        chunk = &ctx->plan_chunks.back();
    }
    char *host = chunk->host + chunk->used, *dev = chunk->dev + chunk->used;
    chunk->used += bytes;
    memset(host, 0, bytes);
    memcpy(host, tasks.data(), tasks.size() * sizeof(int4));
    memcpy(host + off_ptr, dep_ptr.data(), dep_ptr.size() * sizeof(int));
    if (!dep_idx.empty()) memcpy(host + off_idx, dep_idx.data(), dep_idx.size() * sizeof(int));
    FDS_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    plan.tasks = dev;
    plan.dep_ptr = (int *)(dev + off_ptr);
    plan.dep_idx = (int *)(dev + off_idx);
    plan.done = (unsigned *)(dev + off_done);
    ctx->chain_tasks = nullptr;
    plan.n_tasks = (int)tasks.size();
    plan.k = k;
    if (getenv("FDS_DEBUG_PLAN")) {
        long long hmin = rows, hmax = 0;
        for (auto &it : items) { hmin = std::min(hmin, it.ye - it.ys); hmax = std::max(hmax, it.ye - it.ys); }
        int general = 0;
        for (int s = 0; s < n_strips; ++s) general += has_general[(size_t)s];
        fprintf(stderr, "[fds plan] rows %lld strips %d (%d with general rows) k %d: %d tasks, heights %lld..%lld\n",
                rows, n_strips, general, k, plan.n_tasks, hmin, hmax);
    }
    return 0;
}

// Launches a streaming kernel; `overlap`: with programmatic stream serialisation, i.e. its CTAs may
// become resident while the previous kernel on the stream (the previous sweep over the same task
// table) is still running -- the tasks order themselves (TaskSync).
template <typename Kernel, typename Args>
int launch_sweep(fds_ctx *ctx, Kernel kernel, long long ctas, int smem, const Args &args, bool overlap) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kStreamWarps * 32);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    if (ctx->dry_run) {   // make sure the kernel is loaded on this device (lazy module loading)
        cudaFuncAttributes loaded;
        FDS_CUDA(ctx, cudaFuncGetAttributes(&loaded, kernel));
        return 0;
    }
    FDS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, args));
    return 0;
}

template <int K, bool THERMAL, bool STATS, bool AXI>
int launch_stream2d_as(fds_ctx *ctx, const Stream2DArgs &a) {
    auto kernel = stream2d_kernel<K, THERMAL, STATS, AXI>;
    const int smem = kStreamWarps * kS2WarpRingBytes;
    static bool configured_by_device[64] = {};     // function attributes are per device
    bool &configured = configured_by_device[ctx->d.device & 63];
    if (!configured) {
        FDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    // persistent CTAs pull tasks from a counter
    const long long want = (a.n_tasks + kStreamWarps - 1) / kStreamWarps;
    const long long ctas = std::min<long long>(want, 148 * kS2CtasPerSm);
    return launch_sweep(ctx, kernel, ctas, smem, a, a.sync.wait_deps != 0);
}

template <int K, bool THERMAL>
int launch_stream2d(fds_ctx *ctx, const Stream2DArgs &a) {
    if (ctx->axi)
        return a.stats ? launch_stream2d_as<K, THERMAL, true, true>(ctx, a)
                       : launch_stream2d_as<K, THERMAL, false, true>(ctx, a);
    return a.stats ? launch_stream2d_as<K, THERMAL, true, false>(ctx, a)
                   : launch_stream2d_as<K, THERMAL, false, false>(ctx, a);
}

// Resident CTAs per SM, measured on B200 at K = 2 in Gcell-updates/s (lossy Acoustic2D 4096^2 /
// lossy Acoustic3DAxi 8192 x 4096; profiles/r2_c18_ab.jsonl): 3 CTAs (12 warps, 168 registers, spills
// in and around the steady loop) 161 / 106, 2 CTAs (8 warps, 188 / 204 registers, no spills)
// 160 / 121; 4 CTAs (128 registers) 122 / 69 (round 1). The axisymmetric body holds 14 more
// per-column coefficients and the quotient's operands: it runs with 2 CTAs, the plain one with 3.
template <int K, bool AXI, bool VISC, int CTAS>
int launch_streamv_ctas(fds_ctx *ctx, const StreamVArgs &av) {
    auto kernel = streamv_kernel<K, AXI, VISC, CTAS>;
    const int smem = kStreamWarps * kS2WarpRingBytes;
    static bool configured_by_device[64] = {};     // function attributes are per device
    bool &configured = configured_by_device[ctx->d.device & 63];
    if (!configured) {
        FDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // the ring plus the per-column coefficient slots of three CTAs: ~190 KB of the SM's 228
        FDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    const long long want = (av.base.n_tasks + kStreamWarps - 1) / kStreamWarps;
    const long long ctas = std::min<long long>(want, 148 * CTAS);
    return launch_sweep(ctx, kernel, ctas, smem, av, av.base.sync.wait_deps != 0);
}

template <int K, bool AXI, bool VISC>
int launch_streamv(fds_ctx *ctx, const StreamVArgs &av) {
    if constexpr (K == 1) {   // the odd last step of a run
        return launch_streamv_ctas<K, AXI, VISC, kSVCtasPerSm>(ctx, av);
    } else {
        return streamv_ctas_per_sm(ctx) == 2 ? launch_streamv_ctas<K, AXI, VISC, 2>(ctx, av)
                                : launch_streamv_ctas<K, AXI, VISC, 3>(ctx, av);
    }
}

int dispatch_stream2d(fds_ctx *ctx, Stream2DArgs a, int k) {
    const long long rows = a.row_end - a.row_begin;
    if (rows <= 0) return 0;
    const int stride = kS2StripStride;
    a.n_strips = (int)((a.nx + stride - 1) / stride);
    if (!ctx->census_valid && strip_census(ctx, a.n_strips)) return 1;
    fds_ctx::StreamPlan *plan = nullptr;
    for (auto &p : ctx->plans)
        if (p.row_begin == a.row_begin && p.row_end == a.row_end && p.k == k) plan = &p;
    if (!plan) {
        fds_ctx::StreamPlan fresh;
        fresh.row_begin = a.row_begin;
        fresh.row_end = a.row_end;
        // rows a task streams in before its first owned row comes out: k, or 2k with the
        // three-row window of the viscous kernel
        if (build_stream_plan(ctx, fresh, a.n_strips, k, ctx->use_streamv ? 2 * k : k)) return 1;
        ctx->plans.push_back(fresh);
        plan = &ctx->plans.back();
    }
    a.tasks = (const int4 *)plan->tasks;
    a.n_tasks = plan->n_tasks;
    a.stats = ctx->stream_stats;
    a.task_counter = next_counter(ctx);
    if (!a.task_counter) return fail(ctx, "stream2d: counter reset failed");
    // ordering against the previous sweep: a.sync.seq and the halo fields were set by the caller
    a.sync.dep_ptr = plan->dep_ptr;
    a.sync.dep_idx = plan->dep_idx;
    a.sync.done = plan->done;
    a.sync.wait_deps = ctx->overlap && ctx->chain_tasks == plan->tasks;
    a.sync.n_edge_tasks[0] = plan->n_edge[0];
    a.sync.n_edge_tasks[1] = plan->n_edge[1];
    ctx->chain_tasks = ctx->dry_run ? nullptr : plan->tasks;
    a.map = ctx->map + ctx->pad + ctx->halo;
    a.tab = ctx->tab;
    a.tables = ctx->d_tables;
    a.ctab = ctx->ctab;
    a.cvec = ctx->cvec;
    a.n_mat1 = ctx->d.n_materials + 1;
    if (ctx->use_streamv) {
        StreamVArgs av{};
        av.base = a;
        av.ctab = ctx->ctab;
        av.cvec = ctx->cvec;
        av.n_mat1 = ctx->d.n_materials + 1;
        const bool lossy = ctx->d.lossy != 0;
        const bool axi = ctx->d.model == FDS_ACOUSTIC3DAXI;
        if (!lossy || (!axi && ctx->d.model != FDS_ACOUSTIC2D))
            return fail(ctx, "streamv: unsupported model");
#define FDS_STREAMV_CASE(K_)                                                                 \
    case K_:                                                                                 \
        return axi ? launch_streamv<K_, true, true>(ctx, av)                                 \
                   : launch_streamv<K_, false, true>(ctx, av);
        switch (k) {
            FDS_STREAMV_CASE(1)
            FDS_STREAMV_CASE(2)
        }
#undef FDS_STREAMV_CASE
        return fail(ctx, "streamv: bad step count");
    }
#define FDS_STREAM_CASE(K_)                                                                  \
    case K_:                                                                                 \
        return ctx->thermal ? launch_stream2d<K_, true>(ctx, a) : launch_stream2d<K_, false>(ctx, a);
    switch (k) {
        FDS_STREAM_CASE(1)
        FDS_STREAM_CASE(2)
        FDS_STREAM_CASE(3)
        FDS_STREAM_CASE(4)
    }
#undef FDS_STREAM_CASE
    return fail(ctx, "stream2d: bad step count");
}

// Launch with programmatic stream serialisation (the kernels wait for their predecessor themselves).
template <typename Kernel, typename... Args>
cudaError_t launch_chained(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                           Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("FDS_NO_OVERLAP") ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// 1-D tiling: owned cells per CTA, halo and steps per launch.
struct Plan1D {
    int tile, halo, steps, threads, per;
    long long ctas;
    size_t smem;
};

Plan1D plan_1d(const fds_ctx *ctx, long long steps_left) {
    Plan1D p{};
    const long long n = ctx->d.nx;
    if (ctx->d.kernel != 1) {
        // register-resident warp tiles (fds_line1d.cuh): 256 cells per warp, one warp per CTA
        p.per = 0;
        p.threads = 32;
        if (n + 16 <= kLineWidth) {   // the whole line in one warp: any number of steps per launch
            p.tile = (int)n;
            p.halo = 8;
            p.ctas = 1;
            p.steps = (int)std::min<long long>(steps_left, 1 << 20);
        } else {
            // steps per launch = halo / 2: a wide halo saves launches (a launch costs ~7 us of start-up
            // and tail) but leaves fewer owned cells per warp; short lines can afford it
            p.halo = n <= 592 * 64 ? 96 : n <= 1184 * 128 ? 64 : 32;
            if (ctx->halo_1d > 0) p.halo = std::min(112, (ctx->halo_1d + 7) / 8 * 8);
            // a launch that can only advance a few steps (coupled fields: one) needs no wider halo
            p.halo = (int)std::min<long long>(p.halo, (2 * steps_left + 7) / 8 * 8);
            p.tile = kLineWidth - 2 * p.halo;
            p.ctas = (n + p.tile - 1) / p.tile;
            p.steps = (int)std::min<long long>(steps_left, p.halo / 2);
        }
        p.smem = 0;
        return p;
    }
    if (n + 4 <= k1DThreads) {  // a short line in one CTA: no neighbours, any number of steps per launch
        p.tile = (int)n;
        p.halo = 2;
        p.ctas = 1;
        p.steps = (int)std::min<long long>(steps_left, 1 << 20);
    } else {
        // several CTAs, halo / 2 steps per launch. A step costs issue slots and latency of one SM (four
        // phases between block barriers), so tiles are small: one cell per thread, a dozen warps.
        p.halo = ctx->halo_1d > 0 ? ctx->halo_1d : 64;
        p.tile = 256;
        if (ctx->tile_rows > 0) p.tile = ctx->tile_rows;      // FDS_TILE_ROWS: experiments
        p.tile = (int)std::min<long long>(p.tile, (long long)k1DMaxPerThread * k1DThreads - 2 * p.halo);
        p.ctas = (n + p.tile - 1) / p.tile;
        p.steps = (int)std::min<long long>(steps_left, p.halo / 2);
    }
    const int width = p.tile + 2 * p.halo;
    p.per = width <= k1DThreads ? 1 : k1DMaxPerThread;
    p.threads = p.per == 1 ? (width + 31) / 32 * 32 : k1DThreads;
    p.smem = (size_t)width * (16 + sizeof(map_t)) + 16;
    return p;
}

template <bool THERMAL, bool LOSSY, int PER>
int launch_step1d_as(fds_ctx *ctx, const Step1DArgs &a, const StepTables &t, const Plan1D &p) {
    auto kernel = step1d_kernel<THERMAL, LOSSY, PER>;
    static size_t configured_by_device[64] = {};   // function attributes are per device
    size_t &configured = configured_by_device[ctx->d.device & 63];
    if (configured < p.smem) {
        FDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p.smem));
        configured = p.smem;
    }
    kernel<<<(unsigned)p.ctas, p.threads, p.smem, ctx->stream>>>(a, t);
    FDS_CUDA(ctx, cudaGetLastError());
    return 0;
}

template <bool THERMAL, bool LOSSY>
int launch_step1d(fds_ctx *ctx, const Step1DArgs &a, const StepTables &t, const Plan1D &p) {
    if (p.per == 0) {
        auto kernel = line1d_kernel<THERMAL, LOSSY>;
        const int smem = (int)(kLineWarps * sizeof(LineShared));
        static bool configured_by_device[64] = {};
        bool &configured = configured_by_device[ctx->d.device & 63];
        if (!configured) {
            FDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               smem));
            configured = true;
        }
        // consecutive launches of a long run are chained: the next grid is scheduled while this one
        // runs and waits for it on the device (griddepcontrol.wait at the top of the kernel)
        const unsigned ctas = (unsigned)((p.ctas + kLineWarps - 1) / kLineWarps);
        FDS_CUDA(ctx, launch_chained(kernel, dim3(ctas), dim3(32 * kLineWarps), (size_t)smem,
                                     ctx->stream, a, t, (int)p.ctas));
        return 0;
    }
    return p.per == 1 ? launch_step1d_as<THERMAL, LOSSY, 1>(ctx, a, t, p)
                      : launch_step1d_as<THERMAL, LOSSY, k1DMaxPerThread>(ctx, a, t, p);
}

int ensure_ring(fds_ctx *ctx, long long n_steps) {
    // two halves, each holding `ring_half` step records
    // (without probes nothing is recorded: a call is one chunk, so that the step kernels can advance
    // as many steps per launch as they support)
    long long half = std::max<long long>(1, n_steps);
    if (ctx->n_slots > 0) {
        const long long budget = (32ll << 20) / (ctx->n_slots * 8);
        half = std::max<long long>(1, std::min<long long>(n_steps, std::max<long long>(budget, 64)));
    }
    const size_t bytes = ctx->n_slots > 0 ? (size_t)(2 * half * ctx->n_slots) * 8 : 16;
    if (ctx->ring.bytes < bytes || (ctx->n_slots > 0 && ctx->ring_half < half)) {
        if (ctx->ring.ptr) {
            FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->ring.ptr);
            ctx->device_bytes -= (long long)ctx->ring.bytes;
            ctx->ring = DevArray();
        }
        void *p = nullptr;
        if (dev_alloc(ctx, &p, bytes, true)) return 1;
        ctx->ring.ptr = p;
        ctx->ring.bytes = bytes;
    }
    ctx->ring_half = half;
    return 0;
}

// ---- peer-memory halo path ------------------------------------------------------------------------
// A slab of a multi-GPU run: one process per GPU with an NCCL communicator (fds_comm_init), or several
// contexts of one process wired to each other directly (fds_slab_init + fds_peer_connect).
bool is_multi(const fds_ctx *ctx) { return ctx->world > 1 && (ctx->comm || ctx->local_peers); }

bool peers_ready(const fds_ctx *ctx) {
    if (getenv("FDS_NO_PEER")) return false;
    if (ctx->rank > 0 && !ctx->peer_open[0]) return false;
    if (ctx->rank < ctx->world - 1 && !ctx->peer_open[1]) return false;
    // double2 copies: rows must start on 16-byte boundaries
    return ctx->d.rows >= ctx->d.halo_rows && ctx->d.nx % 2 == 0;
}

// Before launch L: the neighbours' rows of launch L-1 must have landed in my halo rows. The first
// launch of a call is covered by the host-driven exchange at the start of run_steps.
int peer_wait(fds_ctx *ctx, bool wait) {
    if (!wait) return 0;
    const unsigned *lo = ctx->rank > 0 ? ctx->flags + 0 : nullptr;
    const unsigned *hi = ctx->rank < ctx->world - 1 ? ctx->flags + 1 : nullptr;
    halo_wait_kernel<<<1, 1, 0, ctx->stream>>>(lo, hi, ctx->launch_seq - 1);
    FDS_CUDA(ctx, cudaGetLastError());
    return 0;
}

// After launch L: my outermost halo_rows rows of buffer `which` go into the neighbours' halo rows of
// the same buffer, then their flags are released with L.
int peer_push(fds_ctx *ctx, int which) {
    const long long nx = ctx->d.nx, h = ctx->d.halo_rows, rows = ctx->d.rows;
    const long long off = ctx->pad + ctx->halo;
    const int ncomp = ctx->thermal ? 1 : 3;
    HaloPushArgs p{};
    p.count = h * nx;
    p.launch_id = ctx->launch_seq;
    p.blocks_done = ctx->flags + 2;
    int seg = 0;
    for (int side = 0; side < 2; ++side) {
        const bool has = side == 0 ? ctx->rank > 0 : ctx->rank < ctx->world - 1;
        p.flag_out[side] = has ? ctx->peer_flags[side] + (1 - side) : nullptr;
        if (!has) continue;
        for (int c = 0; c < ncomp; ++c, ++seg) {
            double *mine = origin(ctx, which, c);
            double *theirs = (double *)ctx->peer_base[side][which][c] + off;
            if (side == 0) {   // my bottom rows -> lower neighbour's upper halo
                p.src[seg] = mine;
                p.dst[seg] = theirs + ctx->peer_rows[0] * nx;
            } else {           // my top rows -> upper neighbour's lower halo
                p.src[seg] = mine + (rows - h) * nx;
                p.dst[seg] = theirs - h * nx;
            }
        }
    }
    p.n_segments = seg;
    if (seg == 0) return 0;
    if (p.count % 2) return fail(ctx, "peer_push: odd halo size");
    const int blocks = (int)std::min<long long>(32, (p.count / 2 + 255) / 256);
    halo_push_kernel<<<dim3((unsigned)blocks, (unsigned)seg), 256, 0, ctx->stream>>>(p);
    FDS_CUDA(ctx, cudaGetLastError());
    return 0;
}

// The same exchange from inside the streaming kernels (TaskSync in fds_stream2d.cuh): tasks that read
// halo rows wait for flags >= wait_seq (0: nothing to wait for), tasks that own edge rows copy them into
// the neighbours' buffer `which` and the last one releases the neighbours' flags with this sweep's number.
void fill_halo_sync(fds_ctx *ctx, TaskSync &y, int which, unsigned wait_seq, bool push) {
    const long long nx = ctx->d.nx, h = ctx->d.halo_rows;
    const long long off = ctx->pad + ctx->halo;
    y.halo_wait_seq = wait_seq;
    y.halo_push = push ? 1 : 0;
    y.ncomp = ctx->thermal ? 1 : 3;
    y.halo_rows = (int)h;
    y.rows = ctx->d.rows;
    y.edge_done = ctx->flags + kFlagEdgeDone;
    for (int side = 0; side < 2; ++side) {
        const bool has = side == 0 ? ctx->rank > 0 : ctx->rank < ctx->world - 1;
        y.flag_in[side] = has ? ctx->flags + side : nullptr;
        y.flag_out[side] = has ? ctx->peer_flags[side] + (1 - side) : nullptr;
        for (int c = 0; c < 3; ++c) {
            y.peer_dst[side][c] = nullptr;
            if (!has || c >= y.ncomp) continue;
            double *theirs = (double *)ctx->peer_base[side][which][c] + off;
            // my bottom rows -> the lower neighbour's upper halo; my top rows -> the upper one's lower halo
            y.peer_dst[side][c] = side == 0 ? theirs + ctx->peer_rows[0] * nx : theirs - h * nx;
        }
    }
}

int exchange_halos(fds_ctx *ctx, int which);

// Slabs of one process, start of a call: the neighbours' edge rows of the current state (fresh uploads,
// complete on every slab before any of them starts stepping -- the host side sees to that) are copied
// into this slab's halo rows with peer copies on this slab's stream. From then on the rows travel from
// inside the step kernels.
int pull_halos_local(fds_ctx *ctx, int which) {
    const long long nx = ctx->d.nx, h = ctx->d.halo_rows, rows = ctx->d.rows;
    const size_t bytes = (size_t)(h * nx) * 8;
    const int ncomp = ctx->thermal ? 1 : 3;
    for (int side = 0; side < 2; ++side) {
        fds_ctx *other = ctx->local_neighbour[side];
        if (!other) continue;
        for (int c = 0; c < ncomp; ++c) {
            double *mine = origin(ctx, which, c);
            const double *theirs = origin(other, which, c);
            // lower neighbour: its top rows -> my lower halo; upper neighbour: its bottom rows -> my upper
            double *dst = side == 0 ? mine - h * nx : mine + rows * nx;
            const double *src = side == 0 ? theirs + (other->d.rows - h) * nx : theirs;
            FDS_CUDA(ctx, cudaMemcpyPeerAsync(dst, ctx->d.device, src, other->d.device, bytes,
                                              ctx->stream));
        }
    }
    return 0;
}

// After a synchronisation: did a device-side wait (neighbour flag, task of the previous sweep) give up?
int check_device_waits(fds_ctx *ctx) {
    unsigned err = 0;
    FDS_CUDA(ctx, cudaMemcpyAsync(&err, ctx->flags + kFlagError, sizeof(err), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) {
        cudaMemsetAsync(ctx->flags + kFlagError, 0, sizeof(unsigned), ctx->stream);
        return fail(ctx, "a device-side wait timed out (neighbour slab or previous sweep never "
                         "arrived): results of this call are invalid");
    }
    return 0;
}

// ---- medium flow ------------------------------------------------------------------------------------
// True if some row moves after leapfrog step `step` (pyfds/acoustic_flow.py:54: step % period == 0).
bool flow_due(const fds_ctx *ctx, long long step) {
    for (long long period : ctx->flow_unique)
        if (period <= 1 || step % period == 0) return true;
    return false;
}

// Steps that may run back to back starting at `step` before a row has to move: the launch may include
// the step after which the shift is due, but none beyond it.
long long flow_run_length(const fds_ctx *ctx, long long step, long long limit) {
    for (long long j = 0; j < limit; ++j)
        if (flow_due(ctx, step + j)) return j + 1;
    return limit;
}

// Moves the rows that are due after `step` in buffer `which` (owned rows; halo rows are refreshed by
// the exchange that follows).
int launch_flow(fds_ctx *ctx, int which, long long step) {
    FlowArgs a{};
    for (int c = 0; c < 3; ++c) a.state[c] = origin(ctx, which, c);
    a.periods = (const long long *)ctx->flow_periods.ptr;
    a.nx = ctx->d.nx;
    a.rows = ctx->d.rows;
    a.step = step;
    flow_shift_kernel<<<dim3((unsigned)ctx->d.rows, 3), kFlowThreads, 0, ctx->stream>>>(a);
    FDS_CUDA(ctx, cudaGetLastError());
    ctx->flow_shifts += 1;
    return 0;
}

// The time loop. `drain` = copy probe records to pinned host memory behind the computation.
int run_steps(fds_ctx *ctx, long long first_step, long long n_steps, bool drain) {
    NvtxRange nvtx_range("fds:step enqueue");
    if (n_steps <= 0) return 0;
    if (!ctx->map_uploaded) return fail(ctx, "fds_step: material map not uploaded");
    if (ctx->n_signals > 0 &&
        (first_step < ctx->sig_first || first_step + n_steps > ctx->sig_first + ctx->sig_steps))
        return fail(ctx, "fds_step: step range outside the uploaded signal window");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (refresh_flags(ctx)) return 1;
    if (ensure_ring(ctx, n_steps)) return 1;
    if (drain && ctx->n_slots > 0) {
        const size_t need = (size_t)n_steps * ctx->n_slots * 8;
        if (ctx->pinned_bytes < need) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);

            ctx->pinned = nullptr;
            ctx->pinned_bytes = 0;
            const size_t alloc_bytes = std::max<size_t>(need, 1u << 20);
            FDS_CUDA(ctx, cudaHostAlloc(&ctx->pinned, alloc_bytes, cudaHostAllocPortable));
            ctx->pinned_bytes = alloc_bytes;
        }
    }

    StepTables t = make_tables(ctx);
    if (ctx->use_stream2d || ctx->use_streamv)
        FDS_CUDA(ctx, cudaMemcpyAsync(ctx->d_tables, &t, sizeof(StepTables), cudaMemcpyHostToDevice,
                                      ctx->stream));
    const long long half = ctx->ring_half;
    const long long sig0 = first_step - ctx->sig_first;
    ctx->last_launches = 0;
    ctx->last_steps_per_launch = 1;
    ctx->flow_shifts = 0;

    if (is_multi(ctx) && ctx->dims == 2) {
        // the neighbours' rows of the current state (fresh upload, or a previous call's last step)
        if (ctx->local_peers) {
            if (pull_halos_local(ctx, ctx->cur)) return 1;
        } else {
            FDS_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
            FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_edge, 0));
            if (exchange_halos(ctx, ctx->cur)) return 1;
            FDS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
            FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
        }
    }
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    ctx->chain_tasks = nullptr;
    long long done = 0;
    long long chunk = 0;
    while (done < n_steps) {
        const long long chunk_steps = std::min(half, n_steps - done);
        const int h = (int)(chunk & 1);
        ctx->chain_tasks = nullptr;   // events sit between the chunks
        if (drain && ctx->n_slots > 0 && chunk >= 2)  // the half must have been drained
            FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_drained[h], 0));
        long long in_chunk = 0;
        while (in_chunk < chunk_steps) {
            const long long s = done + in_chunk;
            const long long ring_row = h * half + in_chunk;
            long long advanced = 1;
            if (ctx->dims == 1) {
                Plan1D p = plan_1d(ctx, chunk_steps - in_chunk);
                Step1DArgs a{};
                a.in[0] = origin(ctx, ctx->cur, 0);
                a.in[1] = origin(ctx, ctx->cur, 1);
                a.out[0] = origin(ctx, ctx->cur ^ 1, 0);
                a.out[1] = origin(ctx, ctx->cur ^ 1, 1);
                a.n = ctx->d.nx;
                a.tile = p.tile;
                a.halo = p.halo;
                a.n_steps = p.steps;
                a.sig_index = sig0 + s;
                a.ring_row = ring_row;
                int rc;
                if (ctx->thermal) rc = launch_step1d<true, false>(ctx, a, t, p);
                else if (ctx->d.lossy) rc = launch_step1d<false, true>(ctx, a, t, p);
                else rc = launch_step1d<false, false>(ctx, a, t, p);
                if (rc) return 1;
                advanced = p.steps;
                if (p.per == 0)
                    ctx->last_kernel = ctx->thermal ? "line1d_kernel<thermal>"
                                       : ctx->d.lossy ? "line1d_kernel<acoustic,lossy>"
                                                      : "line1d_kernel<acoustic,lossless>";
                else
                    ctx->last_kernel = ctx->thermal ? "step1d_kernel<thermal>"
                                       : ctx->d.lossy ? "step1d_kernel<acoustic,lossy>"
                                                      : "step1d_kernel<acoustic,lossless>";
                ctx->last_steps_per_launch = p.steps;
                ctx->last_launches += 1;
            } else if (ctx->use_stream2d || ctx->use_streamv) {
                Stream2DArgs a{};
                for (int c = 0; c < 3; ++c) {
                    a.in[c] = origin(ctx, ctx->cur, c);
                    a.out[c] = origin(ctx, ctx->cur ^ 1, c);
                }
                a.nx = ctx->d.nx;
                a.sig_index = sig0 + s;
                a.ring_row = ring_row;
                const long long rows = ctx->d.rows;
                const bool multi = is_multi(ctx);
                int k = (int)std::min<long long>(std::min(ctx->max_k, stream_max_steps(ctx)),
                                                 chunk_steps - in_chunk);
                // a slab can only advance as many steps as its halo rows cover
                if (multi) k = std::min<int>(k, ctx->d.halo_rows / (ctx->d.lossy ? 2 : 1));
                // rows of a flowing medium move between two steps: end the launch there
                if (ctx->flow) k = (int)flow_run_length(ctx, first_step + s, k);
                const long long last = first_step + s + k - 1;
                const bool shift = ctx->flow && flow_due(ctx, last);
                // thermal fluxes are derived data: stored only by the launch that ends the call
                a.write_vector = (s + k == n_steps);
                a.sync.seq = ++ctx->launch_seq;
                a.sync.error = ctx->flags + kFlagError;
                if (multi && peers_ready(ctx) && ctx->halo_in_kernel) {
                    // One launch for the whole slab; halo rows travel over NVLink peer memory, pushed
                    // and awaited by the tasks themselves. Rows of a flowing medium move after the
                    // sweep and before they travel: that launch pushes with a kernel of its own.
                    const bool first = !(in_chunk > 0 || chunk > 0);
                    fill_halo_sync(ctx, a.sync, ctx->cur ^ 1, first ? 0u : a.sync.seq - 1u, !shift);
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    if (shift) {
                        if (launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                        if (peer_push(ctx, ctx->cur ^ 1)) return 1;
                        ctx->chain_tasks = nullptr;
                    }
                    ctx->last_launches += 1;
                } else if (multi && peers_ready(ctx)) {
                    // the same with a waiting and a pushing kernel either side of the sweep
                    if (peer_wait(ctx, in_chunk > 0 || chunk > 0)) return 1;
                    ctx->chain_tasks = nullptr;
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    if (shift && launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                    if (peer_push(ctx, ctx->cur ^ 1)) return 1;
                    ctx->chain_tasks = nullptr;
                    ctx->last_launches += 1;
                } else if (multi && shift) {
                    // the edge rows move before they travel: no overlap for this launch
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    if (launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_edge, 0));
                    if (exchange_halos(ctx, ctx->cur ^ 1)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
                    ctx->chain_tasks = nullptr;
                    ctx->last_launches += 1;
                } else if (multi) {
                    // the k outermost rows of either side travel while the interior is computed
                    const long long band = std::min<long long>(ctx->d.halo_rows, rows);
                    a.row_begin = 0; a.row_end = band;
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    a.row_begin = std::max(band, rows - band); a.row_end = rows;
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_edge, 0));
                    if (exchange_halos(ctx, ctx->cur ^ 1)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
                    a.row_begin = band; a.row_end = std::max(band, rows - band);
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
                    ctx->chain_tasks = nullptr;
                    ctx->last_launches += 3;
                } else {
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_stream2d(ctx, a, k)) return 1;
                    if (shift) {
                        if (launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                        ctx->chain_tasks = nullptr;
                    }
                    ctx->last_launches += 1;
                }
                advanced = k;
                ctx->last_steps_per_launch = std::max<long long>(ctx->last_steps_per_launch, k);
                ctx->last_kernel =
                    ctx->use_streamv
                        ? (ctx->d.model == FDS_ACOUSTIC3DAXI ? "streamv_kernel<acoustic3daxi,lossy>"
                                                             : "streamv_kernel<acoustic2d,lossy>")
                        : ctx->axi
                              ? (ctx->thermal ? "stream2d_kernel<thermal3daxi>"
                                              : "stream2d_kernel<acoustic3daxi,lossless>")
                              : (ctx->thermal ? "stream2d_kernel<thermal2d>"
                                              : "stream2d_kernel<acoustic2d,lossless>");
            } else {
                Step2DArgs a{};
                for (int c = 0; c < 3; ++c) {
                    a.in[c] = origin(ctx, ctx->cur, c);
                    a.out[c] = origin(ctx, ctx->cur ^ 1, c);
                }
                a.nx = ctx->d.nx;
                a.sig_index = sig0 + s;
                a.ring_row = ring_row;
                // thermal fluxes are derived data: store them only with the last step of the call
                a.write_vector = (s == n_steps - 1);
                const long long rows = ctx->d.rows;
                const long long last = first_step + s;
                const bool shift = ctx->flow && flow_due(ctx, last);
                if (is_multi(ctx) && peers_ready(ctx)) {
                    ++ctx->launch_seq;
                    if (peer_wait(ctx, in_chunk > 0 || chunk > 0)) return 1;
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_step2d(ctx, a, t)) return 1;
                    if (shift && launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                    if (peer_push(ctx, ctx->cur ^ 1)) return 1;
                    ctx->last_launches += 1;
                } else if (is_multi(ctx) && shift) {
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_step2d(ctx, a, t)) return 1;
                    if (launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_edge, 0));
                    if (exchange_halos(ctx, ctx->cur ^ 1)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
                    ctx->last_launches += 1;
                } else if (is_multi(ctx)) {
                    // edge bands first, so that their rows can travel while the interior computes
                    const long long band = std::min<long long>(ctx->d.halo_rows, rows);
                    a.row_begin = 0; a.row_end = band;
                    if (dispatch_step2d(ctx, a, t)) return 1;
                    a.row_begin = std::max(band, rows - band); a.row_end = rows;
                    if (dispatch_step2d(ctx, a, t)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_edge, 0));
                    if (exchange_halos(ctx, ctx->cur ^ 1)) return 1;
                    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
                    a.row_begin = band; a.row_end = std::max(band, rows - band);
                    if (dispatch_step2d(ctx, a, t)) return 1;
                    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
                    ctx->last_launches += 3;
                } else {
                    a.row_begin = 0; a.row_end = rows;
                    if (dispatch_step2d(ctx, a, t)) return 1;
                    if (shift && launch_flow(ctx, ctx->cur ^ 1, last)) return 1;
                    ctx->last_launches += 1;
                }
                ctx->last_kernel = step2d_name(ctx);
            }
            ctx->cur ^= 1;
            in_chunk += advanced;
        }
        if (drain && ctx->n_slots > 0) {
            FDS_CUDA(ctx, cudaEventRecord(ctx->ev_half[h], ctx->stream));
            FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->drain, ctx->ev_half[h], 0));
            FDS_CUDA(ctx, cudaMemcpyAsync((double *)ctx->pinned + done * ctx->n_slots,
                                          (double *)ctx->ring.ptr + h * half * ctx->n_slots,
                                          (size_t)chunk_steps * ctx->n_slots * 8,
                                          cudaMemcpyDeviceToHost, ctx->drain));
            FDS_CUDA(ctx, cudaEventRecord(ctx->ev_drained[h], ctx->drain));
        }
        done += chunk_steps;
        ++chunk;
    }
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    ctx->timed = true;
    return 0;
}

// Sends the outermost owned rows of buffer `which` to the neighbour slabs and receives their rows
// into the halo rows of the same buffer.
int exchange_halos(fds_ctx *ctx, int which) {
    const long long nx = ctx->d.nx, h = ctx->d.halo_rows, rows = ctx->d.rows;
    const size_t count = (size_t)(h * nx);
    const int ncomp = ctx->thermal ? 1 : 3;  // thermal fluxes are recomputed from T
    FDS_NCCL(ctx, g_nccl.GroupStart());
    for (int c = 0; c < ncomp; ++c) {
        double *o = origin(ctx, which, c);
        if (ctx->rank > 0) {
            FDS_NCCL(ctx, g_nccl.Send(o, count, kNcclFloat64, ctx->rank - 1, ctx->comm,
                                      ctx->comm_stream));
            FDS_NCCL(ctx, g_nccl.Recv(o - h * nx, count, kNcclFloat64, ctx->rank - 1, ctx->comm,
                                      ctx->comm_stream));
        }
        if (ctx->rank < ctx->world - 1) {
            FDS_NCCL(ctx, g_nccl.Send(o + (rows - h) * nx, count, kNcclFloat64, ctx->rank + 1,
                                      ctx->comm, ctx->comm_stream));
            FDS_NCCL(ctx, g_nccl.Recv(o + rows * nx, count, kNcclFloat64, ctx->rank + 1, ctx->comm,
                                      ctx->comm_stream));
        }
    }
    FDS_NCCL(ctx, g_nccl.GroupEnd());
    return 0;
}

}  // namespace


namespace {

// ---- Field.simulate in one call: upload, steps and download overlapped by row bands ---------------
// A call that advances few steps is bound by the two PCIe transfers of the state (4096^2: 7 ms up,
// 1 ms of stepping, 7 ms down). The grid is cut into bands of rows; band j is uploaded on one stream
// while the compute stream advances what the rows uploaded so far allow -- launch l (k_l steps, S_l
// steps done after it) may produce the rows below  r1_j - reach * S_l  once band j (rows below r1_j)
// has arrived, the dependency cone of a step being `reach` rows -- and a third stream downloads the
// rows that have reached the final level. PCIe is full duplex, so the call takes about as long as the
// slower of the two transfers instead of their sum. Launches run over row ranges of the same
// ping-pong buffers: launch l reads buffer (cur + l) & 1 and writes the other one; a short first
// launch (n_steps not a multiple of the kernel's steps per launch) keeps k_l non-decreasing, which is
// what keeps the rows a launch overwrites clear of the rows an earlier launch of the NEXT band still
// has to read.
bool pipeline_applies(const fds_ctx *ctx, long long n_steps) {
    if (getenv("FDS_NO_PIPELINE")) return false;
    if (ctx->dims != 2 || !(ctx->use_stream2d || ctx->use_streamv)) return false;
    if (is_multi(ctx) || ctx->flow) return false;
    const int k = std::min(ctx->max_k, stream_max_steps(ctx));
    const long long launches = (n_steps + k - 1) / k;
    if (launches > 16) return false;                   // long runs: the transfers do not matter
    const long long reach = ctx->d.lossy ? 2 : 1;
    // FDS_PIPELINE_FORCE (tests, sanitizer runs): also on grids too small for it to pay
    if (getenv("FDS_PIPELINE_FORCE")) return ctx->d.rows >= 2 * (reach * n_steps + 64);
    // worth it from ~100 MB per component and as long as bands stay much taller than what a call's
    // steps eat off their upper end
    return ctx->d.rows >= 8 * (reach * n_steps + 64) && ctx->owned >= (8ll << 20);
}

int simulate_pipelined(fds_ctx *ctx, long long first_step, long long n_steps,
                       const double *const *in, double *const *out) {
    NvtxRange nvtx_range("fds:simulate pipelined");
    const long long rows = ctx->d.rows, nx = ctx->d.nx;
    const long long reach = ctx->d.lossy ? 2 : 1;
    const int kmax = std::min(ctx->max_k, stream_max_steps(ctx));
    std::vector<int> ks;                                // steps per launch, shortest first
    if (n_steps % kmax) ks.push_back((int)(n_steps % kmax));
    for (long long l = 0; l < n_steps / kmax; ++l) ks.push_back(kmax);
    const int L = (int)ks.size();
    std::vector<long long> done_after(L);               // S_l
    for (int l = 0, acc = 0; l < L; ++l) done_after[l] = (acc += ks[l]);

    long long bands = 8;
    if (const char *env = getenv("FDS_PIPELINE_BANDS")) bands = std::max(2, atoi(env));
    const long long band_rows = std::max<long long>((rows + bands - 1) / bands, reach * n_steps + 64);
    bands = (rows + band_rows - 1) / band_rows;
    while ((long long)ctx->band_events.size() < 2 * bands) {
        cudaEvent_t ev;
        FDS_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->band_events.push_back(ev);
    }
    ctx->last_bands = bands;

    if (refresh_flags(ctx)) return 1;
    if (ensure_ring(ctx, n_steps)) return 1;
    if (ctx->n_slots > 0 && ctx->ring_half < n_steps)
        return fail(ctx, "fds_simulate: probe ring too small");   // (callers chunk long runs)
    if (ctx->n_slots > 0) {
        const size_t need = (size_t)n_steps * ctx->n_slots * 8;
        if (ctx->pinned_bytes < need) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);
            ctx->pinned = nullptr;
            ctx->pinned_bytes = 0;
            const size_t alloc_bytes = std::max<size_t>(need, 1u << 20);
            FDS_CUDA(ctx, cudaHostAlloc(&ctx->pinned, alloc_bytes, cudaHostAllocPortable));
            ctx->pinned_bytes = alloc_bytes;
        }
    }
    StepTables t = make_tables(ctx);
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->d_tables, &t, sizeof(StepTables), cudaMemcpyHostToDevice,
                                  ctx->stream));
    cudaStream_t up = ctx->comm_stream, down = ctx->drain;
    // nothing of an earlier call may still be in flight on the side streams
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
    FDS_CUDA(ctx, cudaStreamWaitEvent(up, ctx->ev_edge, 0));
    FDS_CUDA(ctx, cudaStreamWaitEvent(down, ctx->ev_edge, 0));

    const int cur = ctx->cur;
    const int n_up = ctx->thermal ? 1 : 3;    // thermal fluxes are derived: never read
    for (long long j = 0; j < bands; ++j) {
        const long long r0 = j * band_rows, r1 = std::min(rows, r0 + band_rows);
        for (int c = 0; c < n_up; ++c)
            FDS_CUDA(ctx, cudaMemcpyAsync(origin(ctx, cur, c) + r0 * nx, in[c] + r0 * nx,
                                          (size_t)((r1 - r0) * nx) * 8, cudaMemcpyHostToDevice, up));
        FDS_CUDA(ctx, cudaEventRecord(ctx->band_events[(size_t)(2 * j)], up));
    }
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    ctx->last_launches = 0;
    ctx->last_steps_per_launch = kmax;
    ctx->flow_shifts = 0;
    ctx->chain_tasks = nullptr;
    const long long sig0 = first_step - ctx->sig_first;
    std::vector<long long> produced(L, 0);              // rows [0, produced[l]) exist at level S_l
    long long downloaded = 0;
    const int final_buffer = (cur + L) & 1;
    for (long long j = 0; j < bands; ++j) {
        const long long r1 = std::min(rows, (j + 1) * band_rows);
        FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->band_events[(size_t)(2 * j)], 0));
        for (int l = 0; l < L; ++l) {
            const long long hi = j == bands - 1 ? rows : r1 - reach * done_after[l];
            const long long lo = produced[l];
            if (hi <= lo) continue;
            Stream2DArgs a{};
            for (int c = 0; c < 3; ++c) {
                a.in[c] = origin(ctx, (cur + l) & 1, c);
                a.out[c] = origin(ctx, (cur + l + 1) & 1, c);
            }
            a.nx = nx;
            const long long before = l ? done_after[l - 1] : 0;
            a.sig_index = sig0 + before;
            a.ring_row = before;
            a.write_vector = l == L - 1;
            a.sync.seq = ++ctx->launch_seq;
            a.sync.error = ctx->flags + kFlagError;
            a.row_begin = lo;
            a.row_end = hi;
            ctx->chain_tasks = nullptr;                 // every launch has a task table of its own
            if (dispatch_stream2d(ctx, a, ks[l])) return 1;
            ctx->last_launches += 1;
            produced[l] = hi;
        }
        const long long fin = produced[L - 1];
        if (fin > downloaded) {
            FDS_CUDA(ctx, cudaEventRecord(ctx->band_events[(size_t)(2 * j + 1)], ctx->stream));
            FDS_CUDA(ctx, cudaStreamWaitEvent(down, ctx->band_events[(size_t)(2 * j + 1)], 0));
            for (int c = 0; c < 3; ++c)
                FDS_CUDA(ctx, cudaMemcpyAsync(out[c] + downloaded * nx,
                                              origin(ctx, final_buffer, c) + downloaded * nx,
                                              (size_t)((fin - downloaded) * nx) * 8,
                                              cudaMemcpyDeviceToHost, down));
            downloaded = fin;
        }
    }
    ctx->chain_tasks = nullptr;
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    ctx->timed = true;
    ctx->cur = final_buffer;
    ctx->last_kernel = ctx->use_streamv
                           ? (ctx->d.model == FDS_ACOUSTIC3DAXI ? "streamv_kernel<acoustic3daxi,lossy>"
                                                                : "streamv_kernel<acoustic2d,lossy>")
                       : ctx->axi ? (ctx->thermal ? "stream2d_kernel<thermal3daxi>"
                                                  : "stream2d_kernel<acoustic3daxi,lossless>")
                                  : (ctx->thermal ? "stream2d_kernel<thermal2d>"
                                                  : "stream2d_kernel<acoustic2d,lossless>");
    if (ctx->n_slots > 0)   // all records exist once the last launch is done (stream order)
        FDS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->ring.ptr, (size_t)n_steps * ctx->n_slots * 8,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
}  // namespace

// =================================================================================================
// coupled fields (SynchronizedFields with built-in interactions)
// =================================================================================================

struct fds_group {
    struct Interaction {
        int kind = 0;                 // 0 linear boundary coupling, 1 viscous heating, 2 material law
        int src_member = 0, src_comp = 0, dst_member = 0, dst_comp = 0;
        int additive = 1, accumulate = 0;
        long long stepping = 1;
        double scale = 0.0;
        double *acc = nullptr;        // [n] sum since the last delivery (accumulate)
        double *aux = nullptr;        // viscous heating: density | g | gain, [3][n]
        double dt = 0.0;
        fds::LawArgs law{};           // material law (device pointers filled in)
        double *law_buffers = nullptr;   // statics [3][n] | last [n] | factors [n]
        unsigned long long *law_max = nullptr;
        int *law_count = nullptr;
    };
    std::vector<fds_ctx *> members;
    std::vector<Interaction> interactions;
    long long n = 0;
    StepTables *d_tables = nullptr;   // two members: their tables, for the one-launch pair kernel
    std::string err;
};

namespace {

int gfail(fds_group *g, const std::string &msg) {
    if (g) g->err = msg;
    else {
        std::lock_guard<std::mutex> lock(g_err_mutex);
        g_create_error = msg;
    }
    return 1;
}

#define FDS_GCUDA(g, call)                                                               \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return gfail(g, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

int group_alloc(fds_group *g, void **p, size_t bytes, const void *host) {
    FDS_GCUDA(g, cudaMalloc(p, std::max<size_t>(bytes, 16)));
    if (host) FDS_GCUDA(g, cudaMemcpy(*p, host, bytes, cudaMemcpyHostToDevice));
    else FDS_GCUDA(g, cudaMemset(*p, 0, std::max<size_t>(bytes, 16)));
    return 0;
}

bool group_member_ok(const fds_group *g, int member, int comp) {
    return member >= 0 && member < (int)g->members.size() && comp >= 0 && comp < 2;
}

// One 1-D step of one member on `stream`, recording probes into ring row `ring_row`.
int group_member_step(fds_ctx *ctx, const StepTables &t, long long sig_index, long long ring_row) {
    Plan1D p = plan_1d(ctx, 1);
    Step1DArgs a{};
    a.in[0] = origin(ctx, ctx->cur, 0);
    a.in[1] = origin(ctx, ctx->cur, 1);
    a.out[0] = origin(ctx, ctx->cur ^ 1, 0);
    a.out[1] = origin(ctx, ctx->cur ^ 1, 1);
    a.n = ctx->d.nx;
    a.tile = p.tile;
    a.halo = p.halo;
    a.n_steps = 1;
    a.sig_index = sig_index;
    a.ring_row = ring_row;
    int rc;
    if (ctx->thermal) rc = launch_step1d<true, false>(ctx, a, t, p);
    else if (ctx->d.lossy) rc = launch_step1d<false, true>(ctx, a, t, p);
    else rc = launch_step1d<false, false>(ctx, a, t, p);
    if (rc) return 1;
    ctx->cur ^= 1;
    ctx->last_kernel = p.per == 0 ? (ctx->thermal ? "line1d_kernel<thermal>"
                                     : ctx->d.lossy ? "line1d_kernel<acoustic,lossy>"
                                                    : "line1d_kernel<acoustic,lossless>")
                                  : "step1d_kernel";
    ctx->last_steps_per_launch = 1;
    ctx->last_launches += 1;
    return 0;
}

// Both members of a two-field group in one launch (line1d_pair_kernel).
int member_kind(const fds_ctx *ctx) { return ctx->thermal ? 2 : ctx->d.lossy ? 1 : 0; }

template <int K0, int K1>
int launch_pair(fds_group *g, const LinePairArgs &p, unsigned ctas, cudaStream_t stream) {
    auto kernel = line1d_pair_kernel<K0, K1>;
    const int smem = (int)(kLineWarps * sizeof(LineShared));
    static bool configured_by_device[64] = {};     // function attributes are per device
    bool &configured = configured_by_device[g->members[0]->d.device & 63];
    if (!configured) {
        FDS_GCUDA(g, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    FDS_GCUDA(g, launch_chained(kernel, dim3(ctas, 2), dim3(32 * kLineWarps), (size_t)smem, stream, p,
                                (const StepTables *)g->d_tables));
    return 0;
}

int group_pair_step(fds_group *g, long long step, long long ring_row, cudaStream_t stream) {
    LinePairArgs p{};
    unsigned ctas = 1;
    for (int m = 0; m < 2; ++m) {
        fds_ctx *ctx = g->members[(size_t)m];
        const Plan1D plan = plan_1d(ctx, 1);
        Step1DArgs &a = p.a[m];
        a.in[0] = origin(ctx, ctx->cur, 0);
        a.in[1] = origin(ctx, ctx->cur, 1);
        a.out[0] = origin(ctx, ctx->cur ^ 1, 0);
        a.out[1] = origin(ctx, ctx->cur ^ 1, 1);
        a.n = ctx->d.nx;
        a.tile = plan.tile;
        a.halo = plan.halo;
        a.n_steps = 1;
        a.sig_index = step - ctx->sig_first;
        a.ring_row = ring_row;
        p.n_tiles[m] = (int)plan.ctas;
        ctas = std::max(ctas, (unsigned)((plan.ctas + kLineWarps - 1) / kLineWarps));
        ctx->cur ^= 1;
        ctx->last_kernel = "line1d_pair_kernel";
        ctx->last_steps_per_launch = 1;
        ctx->last_launches += 1;
    }
    const int k0 = member_kind(g->members[0]), k1 = member_kind(g->members[1]);
#define FDS_PAIR(A, B) if (k0 == A && k1 == B) return launch_pair<A, B>(g, p, ctas, stream);
    FDS_PAIR(0, 0) FDS_PAIR(0, 1) FDS_PAIR(0, 2)
    FDS_PAIR(1, 0) FDS_PAIR(1, 1) FDS_PAIR(1, 2)
    FDS_PAIR(2, 0) FDS_PAIR(2, 1) FDS_PAIR(2, 2)
#undef FDS_PAIR
    return gfail(g, "fds_group_step: unknown member kinds");
}

int group_apply(fds_group *g, fds_group::Interaction &it, long long step, cudaStream_t stream) {
    const long long n = g->n;
    const unsigned blocks = (unsigned)((n + kCoupleThreads - 1) / kCoupleThreads);
    const bool due = it.stepping <= 1 || step % it.stepping == 0;
    if (it.kind == 2) {
        if (!due) return 0;
        fds_ctx *src = g->members[(size_t)it.src_member];
        LawArgs a = it.law;
        a.source = origin(src, src->cur, it.src_comp);
        a.cell_tab = g->members[(size_t)it.dst_member]->cell_tab;
        if (a.has_threshold)
            FDS_GCUDA(g, cudaMemsetAsync(a.max_bits, 0, sizeof(unsigned long long), stream));
        FDS_GCUDA(g, launch_chained(couple_law_factors_kernel, dim3(blocks), dim3(kCoupleThreads), 0,
                                    stream, a));
        FDS_GCUDA(g, launch_chained(couple_law_assemble_kernel, dim3(blocks), dim3(kCoupleThreads), 0,
                                    stream, a));
        return 0;
    }
    if (!due && !it.accumulate) return 0;
    fds_ctx *src = g->members[(size_t)it.src_member], *dst = g->members[(size_t)it.dst_member];
    DeliverArgs d{};
    d.target = origin(dst, dst->cur, it.dst_comp);
    d.acc = it.accumulate ? it.acc : nullptr;
    d.n = n;
    d.additive = it.additive;
    d.deliver = due ? 1 : 0;
    if (it.kind == 0) {
        FDS_GCUDA(g, launch_chained(couple_linear_kernel, dim3(blocks), dim3(kCoupleThreads), 0, stream,
                                    d, (const double *)origin(src, src->cur, it.src_comp), it.scale));
    } else {
        HeatingArgs h{};
        h.velocity = origin(src, src->cur, it.src_comp);
        h.density = it.aux;
        h.g = it.aux + n;
        h.gain = it.aux + 2 * n;
        h.dt = it.dt;
        FDS_GCUDA(g, launch_chained(couple_viscous_heating_kernel, dim3(blocks), dim3(kCoupleThreads), 0,
                                    stream, d, h));
    }
    return 0;
}

}  // namespace

// =================================================================================================
// exported entry points
// =================================================================================================

namespace {
// FDS_DEBUG_SEGV=1: print a native backtrace (module + offset, resolvable with addr2line) on SIGSEGV
void segv_handler(int sig, siginfo_t *si, void *) {
    {
        char line[128];
        const int len = snprintf(line, sizeof(line), "[fds segv] fault address %p\n", si->si_addr);
        if (len > 0) (void)!write(2, line, (size_t)len);
    }
    void *frames[64];
    const int n = backtrace(frames, 64);
    for (int k = 0; k < n; ++k) {
        Dl_info info;
        if (dladdr(frames[k], &info) && info.dli_fname) {
            char line[512];
            const int len = snprintf(line, sizeof(line), "[fds segv] %s +0x%lx %s\n", info.dli_fname,
                                     (unsigned long)((char *)frames[k] - (char *)info.dli_fbase),
                                     info.dli_sname ? info.dli_sname : "?");
            if (len > 0) (void)!write(2, line, (size_t)len);
        }
    }
    signal(sig, SIG_DFL);
    raise(sig);
}
struct SegvInstaller {
    SegvInstaller() {
        if (getenv("FDS_DEBUG_SEGV")) {
            struct sigaction sa;
            memset(&sa, 0, sizeof(sa));
            sa.sa_sigaction = segv_handler;
            sa.sa_flags = SA_SIGINFO;
            sigaction(SIGSEGV, &sa, nullptr);
        }
    }
} g_segv_installer;
}  // namespace

extern "C" {

int fds_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *fds_last_error(const fds_ctx *ctx) {
    if (ctx) return ctx->err.c_str();
    return g_create_error.c_str();
}

int fds_create(const fds_desc *desc, fds_ctx **out) {
    if (!desc || !out) return fail(nullptr, "fds_create: null argument");
    *out = nullptr;
    const fds_desc &d = *desc;
    if (d.model < FDS_ACOUSTIC1D || d.model > FDS_THERMAL3DAXI)
        return fail(nullptr, "fds_create: unknown model");
    const bool one_d = (d.model == FDS_ACOUSTIC1D || d.model == FDS_THERMAL1D);
    if (d.nx < 1 || d.ny < 1 || d.rows < 1 || d.row0 < 0 || d.row0 + d.rows > d.ny)
        return fail(nullptr, "fds_create: bad grid or slab extents");
    if (one_d && (d.ny != 1 || d.halo_rows != 0))
        return fail(nullptr, "fds_create: 1-D models are single rows without halo");
    if (!one_d && d.nx < 2)
        return fail(nullptr, "fds_create: 2-D models need at least two samples along x");
    if (d.n_materials < 1 || d.n_materials >= kMaxMaterials)
        return fail(nullptr, "fds_create: 1..31 distinct materials supported");
    if (d.halo_rows < 0 || d.halo_rows > 64)
        return fail(nullptr, "fds_create: halo_rows out of range");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, std::string("fds_create: no CUDA device (") +
                                 (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                                 "); this engine has no CPU path");
    }
    if (d.device < 0 || d.device >= n_dev) return fail(nullptr, "fds_create: bad device ordinal");

    fds_ctx *ctx = new fds_ctx();
    ctx->d = d;
    ctx->dims = one_d ? 1 : 2;
    ctx->ncomp = one_d ? 2 : 3;
    ctx->thermal = (d.model == FDS_THERMAL1D || d.model == FDS_THERMAL2D ||
                    d.model == FDS_THERMAL3DAXI);
    ctx->axi = (d.model == FDS_ACOUSTIC3DAXI || d.model == FDS_THERMAL3DAXI);
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < kMaxClasses; ++k) ctx->cls_signal[c][k] = -1;
    ctx->owned = d.rows * d.nx;
    ctx->halo = (long long)d.halo_rows * d.nx;
    if (d.kernel == 2 && !stream_supported(d) && !streamv_supported(d)) {
        delete ctx;
        return fail(nullptr, "fds_create: the streaming kernels need Acoustic2D, Acoustic3DAxi or "
                             "Thermal2D with nx % 4 == 0 and nx >= 128");
    }
    if (const char *env = getenv("FDS_TILE_ROWS")) ctx->tile_rows = atoi(env);
    ctx->use_stream2d = (d.kernel == 0 || d.kernel == 2) && stream_supported(d);
    ctx->use_streamv = (d.kernel == 0 || d.kernel == 2) && streamv_supported(d);
    ctx->use_tile2d = !one_d && !ctx->use_stream2d && !ctx->use_streamv &&
                      (d.kernel == 0 || d.kernel == 3) && d.nx % 8 == 0 && d.nx >= kTileW;
    if (d.kernel == 3 && !ctx->use_tile2d) {
        delete ctx;
        return fail(nullptr, "fds_create: the tile kernel needs a 2-D model with nx % 8 == 0, nx >= 128");
    }
    if (const char *env = getenv("FDS_TILE_ROWS")) ctx->tile_rows = atoi(env);
    if (const char *env = getenv("FDS_CHUNK_ROWS")) ctx->chunk_rows = atoi(env);
    if (const char *env = getenv("FDS_HALO_1D")) ctx->halo_1d = std::max(2, atoi(env) / 2 * 2);
    if (getenv("FDS_NO_OVERLAP")) ctx->overlap = false;
    if (const char *env = getenv("FDS_SV_CTAS")) ctx->sv_ctas = atoi(env) == 2 ? 2 : 3;
    if (getenv("FDS_HALO_KERNELS")) ctx->halo_in_kernel = false;
    if (const char *env = getenv("FDS_MAX_K"))
        ctx->max_k = std::max(1, std::min(kMaxStreamSteps, atoi(env)));
    // >= 9 rows + 4096 cells of padding, rounded so that the origin stays 512-byte aligned
    long long pad = (one_d ? 0 : 9 * d.nx) + 4096;
    pad = (pad + ctx->halo + 63) / 64 * 64 - ctx->halo;
    while ((pad + ctx->halo) % 64) ++pad;
    ctx->pad = pad;
    ctx->alloc = ctx->owned + 2 * (ctx->halo + ctx->pad);

    auto bail = [&](int) {
        std::string msg = ctx->err;
        fds_destroy(ctx);
        return fail(nullptr, "fds_create: " + msg);
    };
#define FDS_TRY(expr) if (expr) return bail(0)
    if (cudaSetDevice(d.device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(0); }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->drain, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) != cudaSuccess) {
        ctx->err = "cudaStreamCreate failed";
        return bail(0);
    }
    cudaEvent_t *plain[] = {&ctx->ev_half[0], &ctx->ev_half[1], &ctx->ev_drained[0],
                            &ctx->ev_drained[1], &ctx->ev_edge, &ctx->ev_comm};
    for (cudaEvent_t *ev : plain)
        if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) {
            ctx->err = "cudaEventCreate failed";
            return bail(0);
        }
    if (cudaEventCreate(&ctx->ev_t0) != cudaSuccess || cudaEventCreate(&ctx->ev_t1) != cudaSuccess) {
        ctx->err = "cudaEventCreate failed";
        return bail(0);
    }
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < ctx->ncomp; ++c)
            FDS_TRY(dev_alloc(ctx, (void **)&ctx->buf[b][c], (size_t)ctx->alloc * 8, true));
    FDS_TRY(dev_alloc(ctx, (void **)&ctx->map, (size_t)ctx->alloc * sizeof(map_t), true));
    FDS_TRY(dev_alloc(ctx, (void **)&ctx->tab, sizeof(double) * FDS_TAB_COUNT * kMaxMaterials, true));
    if (ctx->axi) {
        FDS_TRY(dev_alloc(ctx, (void **)&ctx->ctab,
                          sizeof(double) * FDS_CTAB_COUNT * (d.n_materials + 1) * d.nx, true));
        FDS_TRY(dev_alloc(ctx, (void **)&ctx->cvec, sizeof(double) * FDS_CVEC_COUNT * d.nx, true));
    }
    FDS_TRY(dev_alloc(ctx, (void **)&ctx->flags, kFlagWords * sizeof(unsigned), true));
    FDS_TRY(dev_alloc(ctx, (void **)&ctx->d_tables, sizeof(StepTables), true));
    FDS_TRY(dev_alloc(ctx, (void **)&ctx->task_counters, sizeof(int) * kCounterPool, true));
    if (getenv("FDS_STREAM_STATS"))
        FDS_TRY(dev_alloc(ctx, (void **)&ctx->stream_stats, sizeof(unsigned long long) * 8, true));
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        ctx->err = "device initialisation failed";
        return bail(0);
    }
#undef FDS_TRY
    *out = ctx;
    return 0;
}

void fds_destroy(fds_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->d.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->drain) cudaStreamSynchronize(ctx->drain);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 3; ++c)
            if (ctx->buf[b][c]) cudaFree(ctx->buf[b][c]);
    if (ctx->map) cudaFree(ctx->map);
    if (ctx->tab) cudaFree(ctx->tab);
    if (ctx->cell_tab) cudaFree(ctx->cell_tab);
    if (ctx->ctab) cudaFree(ctx->ctab);
    if (ctx->cvec) cudaFree(ctx->cvec);
    for (int side = 0; side < 2; ++side) {
        if (!ctx->peer_open[side] || ctx->local_peers) continue;
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 3; ++c)
                if (ctx->peer_base[side][b][c]) cudaIpcCloseMemHandle(ctx->peer_base[side][b][c]);
        if (ctx->peer_flags[side]) cudaIpcCloseMemHandle(ctx->peer_flags[side]);
    }
    if (ctx->flags) cudaFree(ctx->flags);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->task_counters) cudaFree(ctx->task_counters);
    if (ctx->stream_stats) cudaFree(ctx->stream_stats);
    for (int c = 0; c < 3; ++c) {
        DevArray *arrays[] = {&ctx->bcells[c], &ctx->boffsets[c], &ctx->balpha[c], &ctx->bvalue[c],
                              &ctx->bsignal[c], &ctx->pcells[c], &ctx->pslots[c],
                              &ctx->browptr[c], &ctx->prowptr[c], &ctx->ccells[c], &ctx->cclass[c]};
        for (DevArray *a : arrays)
            if (a->ptr) cudaFree(a->ptr);
    }
    if (ctx->flagged.ptr) cudaFree(ctx->flagged.ptr);
    if (ctx->line_index.ptr) cudaFree(ctx->line_index.ptr);
    if (ctx->line_entries.ptr) cudaFree(ctx->line_entries.ptr);
    invalidate_plans(ctx);
    for (auto &chunk : ctx->plan_chunks) {
        if (chunk.dev) cudaFree(chunk.dev);
        if (chunk.host) cudaFreeHost(chunk.host);
    }
    ctx->plan_chunks.clear();
    if (ctx->signals.ptr) cudaFree(ctx->signals.ptr);
    if (ctx->ring.ptr) cudaFree(ctx->ring.ptr);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->flow_periods.ptr) cudaFree(ctx->flow_periods.ptr);
    for (int slot = 0; slot < 2; ++slot) {
        if (ctx->frame_dev[slot]) cudaFree(ctx->frame_dev[slot]);
        if (ctx->frame_host[slot]) cudaFreeHost(ctx->frame_host[slot]);
        if (ctx->ev_frame_ready[slot]) cudaEventDestroy(ctx->ev_frame_ready[slot]);
        if (ctx->ev_frame_done[slot]) cudaEventDestroy(ctx->ev_frame_done[slot]);
    }

    for (cudaEvent_t ev : ctx->band_events) cudaEventDestroy(ev);
    cudaEvent_t events[] = {ctx->ev_half[0], ctx->ev_half[1], ctx->ev_drained[0], ctx->ev_drained[1],
                            ctx->ev_t0, ctx->ev_t1, ctx->ev_edge, ctx->ev_comm};
    for (cudaEvent_t ev : events)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->drain) cudaStreamDestroy(ctx->drain);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
}

int fds_upload_material_map(fds_ctx *ctx, const uint8_t *ids, int64_t n) {
    NvtxRange nvtx_range("fds:bake material map h2d");
    if (!ctx || !ids) return fail(ctx, "fds_upload_material_map: null argument");
    if (n != ctx->owned + 2 * ctx->halo)
        return fail(ctx, "fds_upload_material_map: expected (rows + 2*halo_rows) * nx ids");
    for (int64_t k = 0; k < n; ++k)
        if (ids[k] > ctx->d.n_materials)
            return fail(ctx, "fds_upload_material_map: id exceeds n_materials");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    uint8_t *staging = nullptr;
    FDS_CUDA(ctx, cudaMalloc(&staging, (size_t)n));
    cudaError_t e = cudaMemcpyAsync(staging, ids, (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        widen_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->map + ctx->pad,
                                                                              staging, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(staging);
    if (e != cudaSuccess)
        return fail(ctx, std::string("fds_upload_material_map: ") + cudaGetErrorString(e));
    ctx->map_uploaded = true;
    ctx->n_flagged = 0;      // the upload overwrote all flag bits
    ctx->flags_dirty = true;
    return 0;
}

int fds_upload_table(fds_ctx *ctx, int32_t table, const double *values, int64_t n) {
    if (!ctx || !values) return fail(ctx, "fds_upload_table: null argument");
    if (table < 0 || table >= FDS_TAB_COUNT) return fail(ctx, "fds_upload_table: bad table id");
    if (n != ctx->d.n_materials + 1)
        return fail(ctx, "fds_upload_table: expected n_materials + 1 values");
    if (values[0] != 0.0) return fail(ctx, "fds_upload_table: entry 0 (void material) must be 0");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->tab + (size_t)table * kMaxMaterials, values, (size_t)n * 8,
                                  cudaMemcpyHostToDevice, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fds_upload_cell_table(fds_ctx *ctx, int32_t table, const double *values, int64_t n) {
    if (!ctx) return fail(ctx, "fds_upload_cell_table: null context");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (!values) {   // back to the per-material table
        if (ctx->cell_tab) {
            FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->cell_tab);
            ctx->device_bytes -= (long long)sizeof(double) * FDS_TAB_COUNT * ctx->owned;
            ctx->cell_tab = nullptr;
            if (ctx->dims == 2) {   // back to the kernels chosen at creation
                ctx->use_stream2d = ctx->kernels_before_cells[0];
                ctx->use_streamv = ctx->kernels_before_cells[1];
                ctx->use_tile2d = ctx->kernels_before_cells[2];
                invalidate_plans(ctx);
            }
        }
        return 0;
    }
    if (ctx->dims == 1 && ctx->d.kernel == 1)
        return fail(ctx, "fds_upload_cell_table: the shared-memory 1-D kernel reads the material table "
                         "only (use kernel 0)");
    if (ctx->dims == 2 && (ctx->axi || ctx->d.halo_rows != 0 || ctx->d.rows != ctx->d.ny))
        return fail(ctx, "fds_upload_cell_table: 2-D per-cell coefficients need a plain (not "
                         "axisymmetric) model on a single slab");
    if (table < 0 || table >= FDS_TAB_COUNT) return fail(ctx, "fds_upload_cell_table: bad table id");
    if (n != ctx->owned) return fail(ctx, "fds_upload_cell_table: expected one value per cell");
    if (!ctx->cell_tab) {
        if (dev_alloc(ctx, (void **)&ctx->cell_tab, sizeof(double) * FDS_TAB_COUNT * (size_t)n, true))
            return 1;
        if (ctx->dims == 2) {
            // only the one-thread-per-cell kernel reads per-cell coefficients (one step per launch)
            ctx->kernels_before_cells[0] = ctx->use_stream2d;
            ctx->kernels_before_cells[1] = ctx->use_streamv;
            ctx->kernels_before_cells[2] = ctx->use_tile2d;
            ctx->use_stream2d = ctx->use_streamv = ctx->use_tile2d = false;
            invalidate_plans(ctx);
        }
    }
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->cell_tab + (size_t)table * (size_t)n, values,
                                  sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fds_upload_column_table(fds_ctx *ctx, int32_t table, const double *values, int64_t n) {
    if (!ctx || !values) return fail(ctx, "fds_upload_column_table: null argument");
    if (!ctx->axi) return fail(ctx, "fds_upload_column_table: model is not axisymmetric");
    if (table < 0 || table >= FDS_CTAB_COUNT)
        return fail(ctx, "fds_upload_column_table: bad table id");
    const int64_t expect = (int64_t)(ctx->d.n_materials + 1) * ctx->d.nx;
    if (n != expect) return fail(ctx, "fds_upload_column_table: expected (n_materials+1)*nx values");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->ctab + (size_t)table * expect, values, (size_t)n * 8,
                                  cudaMemcpyHostToDevice, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fds_upload_column_vector(fds_ctx *ctx, int32_t vec, const double *values, int64_t n) {
    if (!ctx || !values) return fail(ctx, "fds_upload_column_vector: null argument");
    if (!ctx->axi) return fail(ctx, "fds_upload_column_vector: model is not axisymmetric");
    if (vec < 0 || vec >= FDS_CVEC_COUNT) return fail(ctx, "fds_upload_column_vector: bad id");
    if (n != ctx->d.nx) return fail(ctx, "fds_upload_column_vector: expected nx values");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->cvec + (size_t)vec * ctx->d.nx, values, (size_t)n * 8,
                                  cudaMemcpyHostToDevice, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fds_upload_boundaries(fds_ctx *ctx, int32_t component, const int64_t *cells,
                          const int32_t *offsets, int64_t n_cells, const double *alpha,
                          const double *value, const int32_t *signal, int64_t n_ops) {
    NvtxRange nvtx_range("fds:bake boundary table h2d");
    if (!ctx) return fail(ctx, "fds_upload_boundaries: null context");
    if (component < 0 || component >= ctx->ncomp)
        return fail(ctx, "fds_upload_boundaries: bad component");
    if (n_cells < 0 || n_ops < n_cells || n_cells > 0x7fffffff || n_ops > 0x7fffffff)
        return fail(ctx, "fds_upload_boundaries: bad sizes");
    if (n_cells > 0) {
        if (!cells || !offsets || !alpha || !value || !signal)
            return fail(ctx, "fds_upload_boundaries: null table");
        const long long lo = -ctx->halo, hi = ctx->owned + ctx->halo;
        for (int64_t k = 0; k < n_cells; ++k) {
            if (cells[k] < lo || cells[k] >= hi)
                return fail(ctx, "fds_upload_boundaries: cell outside the slab");
            if (k && cells[k] <= cells[k - 1])
                return fail(ctx, "fds_upload_boundaries: cells must be strictly ascending");
            if (offsets[k + 1] <= offsets[k])
                return fail(ctx, "fds_upload_boundaries: offsets must be strictly ascending");
        }
        if (offsets[0] != 0 || offsets[n_cells] != n_ops)
            return fail(ctx, "fds_upload_boundaries: offsets do not cover the operations");
    }
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    const int c = component;

    // Cells with exactly one operation whose value is a scalar -- or, for the 2-D models, a signal (a
    // source line driven by one signal is a handful of classes at most) -- get a class (applied inline
    // by the kernels, no table access); all others stay in the lookup table of the slow path.
    const bool signal_classes = ctx->dims == 2;
    int n_classes = 1;
    for (int k = 0; k < kMaxClasses; ++k) {
        ctx->cls_alpha[c][k] = ctx->cls_value[c][k] = 0.0;
        ctx->cls_signal[c][k] = -1;
    }
    std::vector<long long> class_cells, slow_cells;
    std::vector<int> class_ids, slow_offsets, slow_signal;
    std::vector<double> slow_alpha, slow_value;
    slow_offsets.push_back(0);
    for (int64_t k = 0; k < n_cells; ++k) {
        const int b = offsets[k], e = offsets[k + 1];
        int cls = 0;
        if (e - b == 1 && (signal[b] < 0 || signal_classes)) {
            const double constant = signal[b] < 0 ? value[b] : 0.0;
            for (int j = 1; j < n_classes && !cls; ++j)
                if (memcmp(&ctx->cls_alpha[c][j], &alpha[b], 8) == 0 &&
                    memcmp(&ctx->cls_value[c][j], &constant, 8) == 0 &&
                    ctx->cls_signal[c][j] == (signal[b] < 0 ? -1 : signal[b]))
                    cls = j;
            if (!cls && n_classes < kMaxClasses) {
                cls = n_classes++;
                ctx->cls_alpha[c][cls] = alpha[b];
                ctx->cls_value[c][cls] = constant;
                ctx->cls_signal[c][cls] = signal[b] < 0 ? -1 : signal[b];
            }
        }
        if (cls) {
            class_cells.push_back(cells[k]);
            class_ids.push_back(cls);
        } else {
            slow_cells.push_back(cells[k]);
            for (int o = b; o < e; ++o) {
                slow_alpha.push_back(alpha[o]);
                slow_value.push_back(value[o]);
                slow_signal.push_back(signal[o]);
            }
            slow_offsets.push_back((int)slow_alpha.size());
        }
    }
    const size_t ns = slow_cells.size(), nops = slow_alpha.size(), nc = class_cells.size();
    if (dev_upload(ctx, ctx->bcells[c], slow_cells.data(), ns * 8)) return 1;
    if (dev_upload(ctx, ctx->boffsets[c], slow_offsets.data(), (ns ? ns + 1 : 0) * 4)) return 1;
    if (dev_upload(ctx, ctx->balpha[c], slow_alpha.data(), nops * 8)) return 1;
    if (dev_upload(ctx, ctx->bvalue[c], slow_value.data(), nops * 8)) return 1;
    if (dev_upload(ctx, ctx->bsignal[c], slow_signal.data(), nops * 4)) return 1;
    if (upload_row_ptr(ctx, ctx->browptr[c], slow_cells.data(), (int64_t)ns)) return 1;
    if (dev_upload(ctx, ctx->ccells[c], class_cells.data(), nc * 8)) return 1;
    if (dev_upload(ctx, ctx->cclass[c], class_ids.data(), nc * 4)) return 1;
    ctx->host_bcells[c] = slow_cells;
    ctx->host_ccells[c] = class_cells;
    if (ctx->dims == 1) {
        ctx->host_boffsets[c] = slow_offsets;
        ctx->host_balpha[c] = slow_alpha;
        ctx->host_bvalue[c] = slow_value;
        ctx->host_bsignal[c] = slow_signal;
    }
    invalidate_plans(ctx);
    ctx->n_bcells[c] = (long long)ns;
    ctx->n_ccells[c] = (long long)nc;
    ctx->flags_dirty = true;
    return 0;
}

int fds_upload_signals(fds_ctx *ctx, const double *samples, int64_t n_signals, int64_t n_steps,
                       int64_t first_step) {
    NvtxRange nvtx_range("fds:bake signals h2d");
    if (!ctx) return fail(ctx, "fds_upload_signals: null context");
    if (n_signals < 0 || n_steps < 0) return fail(ctx, "fds_upload_signals: bad sizes");
    if (n_signals > 0 && n_steps > 0 && !samples)
        return fail(ctx, "fds_upload_signals: null samples");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (dev_upload(ctx, ctx->signals, samples, (size_t)(n_signals * n_steps) * 8)) return 1;
    ctx->n_signals = n_signals;
    ctx->sig_steps = n_steps;
    ctx->sig_first = first_step;
    return 0;
}

int fds_upload_probes(fds_ctx *ctx, int32_t component, const int64_t *cells, const int32_t *slots,
                      int64_t n, int64_t n_slots_total) {
    NvtxRange nvtx_range("fds:bake probe table h2d");
    if (!ctx) return fail(ctx, "fds_upload_probes: null context");
    if (component < 0 || component >= ctx->ncomp) return fail(ctx, "fds_upload_probes: bad component");
    if (n < 0 || n > 0x7fffffff || n_slots_total < 0 || n_slots_total > 0x7fffffff)
        return fail(ctx, "fds_upload_probes: bad sizes");
    for (int64_t k = 0; k < n; ++k) {
        if (cells[k] < 0 || cells[k] >= ctx->owned)
            return fail(ctx, "fds_upload_probes: cell not owned by the slab");
        if (k && cells[k] < cells[k - 1])
            return fail(ctx, "fds_upload_probes: cells must be ascending");
        if (slots[k] < 0 || slots[k] >= n_slots_total)
            return fail(ctx, "fds_upload_probes: slot out of range");
    }
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (dev_upload(ctx, ctx->pcells[component], cells, (size_t)n * 8)) return 1;
    if (dev_upload(ctx, ctx->pslots[component], slots, (size_t)n * 4)) return 1;
    if (upload_row_ptr(ctx, ctx->prowptr[component], (const long long *)cells, n)) return 1;
    ctx->host_pslots[component].assign(slots, slots + n);
    if (ctx->dims == 1) ctx->host_pcells[component].assign(cells, cells + n);
    ctx->n_probes[component] = n;
    ctx->n_slots = n_slots_total;
    ctx->flags_dirty = true;
    return 0;
}

int fds_upload_state(fds_ctx *ctx, int32_t component, const double *values, int64_t n) {
    NvtxRange nvtx_range("fds:state h2d");
    if (!ctx || !values) return fail(ctx, "fds_upload_state: null argument");
    if (component < 0 || component >= ctx->ncomp) return fail(ctx, "fds_upload_state: bad component");
    if (n != ctx->owned) return fail(ctx, "fds_upload_state: expected rows * nx values");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    return state_copy(ctx, origin(ctx, ctx->cur, component), const_cast<double *>(values),
                       (size_t)n * 8, true);
}

int fds_download_state(fds_ctx *ctx, int32_t component, double *values, int64_t n) {
    NvtxRange nvtx_range("fds:state d2h");
    if (!ctx || !values) return fail(ctx, "fds_download_state: null argument");
    if (component < 0 || component >= ctx->ncomp)
        return fail(ctx, "fds_download_state: bad component");
    if (n != ctx->owned) return fail(ctx, "fds_download_state: expected rows * nx values");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    return state_copy(ctx, origin(ctx, ctx->cur, component), values, (size_t)n * 8, false);
}

int fds_host_register(void *host, int64_t bytes) {
    if (!host || bytes <= 0) return fail(nullptr, "fds_host_register: bad argument");
    cudaError_t e = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
    return 0;
}

int fds_host_unregister(void *host) {
    if (!host) return fail(nullptr, "fds_host_unregister: null pointer");
    cudaError_t e = cudaHostUnregister(host);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
    }
    return 0;
}

int fds_reset_state(fds_ctx *ctx) {
    if (!ctx) return fail(ctx, "fds_reset_state: null context");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < ctx->ncomp; ++c)
            FDS_CUDA(ctx, cudaMemsetAsync(ctx->buf[b][c], 0, (size_t)ctx->alloc * 8, ctx->stream));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fds_step(fds_ctx *ctx, int64_t first_step, int64_t n_steps, double *probes_out) {
    NvtxRange nvtx_range("fds:step + probe drain");
    if (!ctx) return fail(ctx, "fds_step: null context");
    if (n_steps < 0) return fail(ctx, "fds_step: negative step count");
    if (ctx->n_slots > 0 && n_steps > 0 && !probes_out)
        return fail(ctx, "fds_step: probes are configured but probes_out is NULL");
    if (run_steps(ctx, first_step, n_steps, true)) return 1;
    if (check_device_waits(ctx)) return 1;
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->drain));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    if (ctx->n_slots > 0 && n_steps > 0) {
        // copy only the slots this slab owns; the others belong to other ranks
        std::vector<char> mine((size_t)ctx->n_slots, 0);
        bool all = true;
        {
            for (int c = 0; c < ctx->ncomp; ++c)
                for (int s : ctx->host_pslots[c]) mine[(size_t)s] = 1;
            for (char m : mine) all = all && m;
        }
        const double *src = (const double *)ctx->pinned;
        if (all) {
            memcpy(probes_out, src, (size_t)n_steps * ctx->n_slots * 8);
        } else {
            for (int64_t s = 0; s < n_steps; ++s)
                for (int64_t k = 0; k < ctx->n_slots; ++k)
                    if (mine[(size_t)k]) probes_out[s * ctx->n_slots + k] = src[s * ctx->n_slots + k];
        }
    }
    return 0;
}

int fds_simulate(fds_ctx *ctx, int64_t first_step, int64_t n_steps, const double *const *values_in,
                 double *const *values_out, double *probes_out) {
    NvtxRange nvtx_range("fds:simulate");
    if (!ctx || !values_in || !values_out) return fail(ctx, "fds_simulate: null argument");
    if (n_steps <= 0) return fail(ctx, "fds_simulate: needs at least one step");
    for (int c = 0; c < ctx->ncomp; ++c)
        if (!values_in[c] || !values_out[c]) return fail(ctx, "fds_simulate: null component array");
    if (ctx->n_slots > 0 && !probes_out)
        return fail(ctx, "fds_simulate: probes are configured but probes_out is NULL");
    if (!ctx->map_uploaded) return fail(ctx, "fds_simulate: material map not uploaded");
    if (ctx->n_signals > 0 &&
        (first_step < ctx->sig_first || first_step + n_steps > ctx->sig_first + ctx->sig_steps))
        return fail(ctx, "fds_simulate: step range outside the uploaded signal window");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    ctx->last_bands = 0;
    bool pipelined = pipeline_applies(ctx, n_steps);
    if (pipelined && ctx->n_slots > 0) {
        // all probe records of the call must fit the device ring (they are drained at the end)
        const long long budget = (32ll << 20) / (ctx->n_slots * 8);
        if (n_steps > std::max<long long>(budget, 64)) pipelined = false;
    }
    if (pipelined) {
        if (simulate_pipelined(ctx, first_step, n_steps, values_in, values_out)) return 1;
    } else {
        // the same three phases back to back: copies of all components enqueued together, one wait
        for (int c = 0; c < ctx->ncomp; ++c)
            FDS_CUDA(ctx, cudaMemcpyAsync(origin(ctx, ctx->cur, c), values_in[c],
                                          (size_t)ctx->owned * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (run_steps(ctx, first_step, n_steps, true)) return 1;
        for (int c = 0; c < ctx->ncomp; ++c)
            FDS_CUDA(ctx, cudaMemcpyAsync(values_out[c], origin(ctx, ctx->cur, c),
                                          (size_t)ctx->owned * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (check_device_waits(ctx)) return 1;
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->drain));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    if (ctx->n_slots > 0)
        memcpy(probes_out, ctx->pinned, (size_t)n_steps * ctx->n_slots * 8);
    return 0;
}

int fds_last_pipeline_bands(fds_ctx *ctx, int64_t *bands) {
    if (!ctx || !bands) return fail(ctx, "fds_last_pipeline_bands: null argument");
    *bands = ctx->last_bands;
    return 0;
}

int fds_step_prepare(fds_ctx *ctx, int64_t first_step, int64_t n_steps) {
    NvtxRange nvtx_range("fds:step prepare");
    if (!ctx) return fail(ctx, "fds_step_prepare: null context");
    if (n_steps <= 0) return 0;
    if (!ctx->map_uploaded) return fail(ctx, "fds_step_prepare: material map not uploaded");
    (void)first_step;
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (refresh_flags(ctx)) return 1;
    if (ensure_ring(ctx, n_steps)) return 1;
    if (ctx->n_slots > 0) {
        const size_t need = (size_t)n_steps * ctx->n_slots * 8;
        if (ctx->pinned_bytes < need) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);
            ctx->pinned = nullptr;
            ctx->pinned_bytes = 0;
            const size_t alloc_bytes = std::max<size_t>(need, 1u << 20);
            FDS_CUDA(ctx, cudaHostAlloc(&ctx->pinned, alloc_bytes, cudaHostAllocPortable));
            ctx->pinned_bytes = alloc_bytes;
        }
    }
    if (ctx->dims == 2) {
        // everything the launches of the call may need: census, task tables for every step count a
        // launch can have, the kernels themselves (lazy module loading), the helpers of the exchange
        ctx->dry_run = true;
        int rc = 0;
        if (ctx->use_stream2d || ctx->use_streamv) {
            const int kmax = std::min(ctx->max_k, stream_max_steps(ctx));
            for (int k = 1; k <= kmax && !rc; ++k) {
                Stream2DArgs a{};
                a.nx = ctx->d.nx;
                a.row_begin = 0;
                a.row_end = ctx->d.rows;
                rc = dispatch_stream2d(ctx, a, k);
            }
        } else {
            Step2DArgs a{};
            a.nx = ctx->d.nx;
            a.row_begin = 0;
            a.row_end = ctx->d.rows;
            rc = dispatch_step2d(ctx, a, make_tables(ctx));
        }
        ctx->dry_run = false;
        ctx->chain_tasks = nullptr;
        if (rc) return 1;
        cudaFuncAttributes loaded;
        FDS_CUDA(ctx, cudaFuncGetAttributes(&loaded, halo_wait_kernel));
        FDS_CUDA(ctx, cudaFuncGetAttributes(&loaded, halo_push_kernel));
        FDS_CUDA(ctx, cudaFuncGetAttributes(&loaded, flow_shift_kernel));
    }
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fds_step_async(fds_ctx *ctx, int64_t first_step, int64_t n_steps) {
    if (!ctx) return fail(ctx, "fds_step_async: null context");
    if (n_steps < 0) return fail(ctx, "fds_step_async: negative step count");
    return run_steps(ctx, first_step, n_steps, false);
}

int fds_sync(fds_ctx *ctx) {
    NvtxRange nvtx_range("fds:sync");
    if (!ctx) return fail(ctx, "fds_sync: null context");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (check_device_waits(ctx)) return 1;
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->drain));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    return 0;
}

int fds_set_flow(fds_ctx *ctx, const int64_t *periods, int64_t n) {
    if (!ctx) return fail(ctx, "fds_set_flow: null context");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (!periods || n == 0) {
        ctx->flow = false;
        ctx->flow_unique.clear();
        return 0;
    }
    if (ctx->d.model != FDS_ACOUSTIC2D)
        return fail(ctx, "fds_set_flow: medium flow is defined for Acoustic2D only");
    // n = ny: the periods of ALL grid rows (every slab of a multi-GPU run passes the same array); the
    // launches of all slabs must end at the same steps, so the schedule (`flow_unique`: where any row
    // of the GRID is due) comes from all of them while the shift kernel only sees the owned rows.
    // n = rows: the owned rows only (single slab).
    if (n != ctx->d.rows && n != ctx->d.ny)
        return fail(ctx, "fds_set_flow: expected one period per grid row (or per owned row)");
    if (n != ctx->d.ny && ctx->d.rows != ctx->d.ny)
        return fail(ctx, "fds_set_flow: a slab of a larger grid needs the periods of all grid rows");
    for (int64_t k = 0; k < n; ++k)
        if (periods[k] < 0) return fail(ctx, "fds_set_flow: periods must not be negative");
    std::vector<long long> unique(periods, periods + n);
    std::sort(unique.begin(), unique.end());
    unique.erase(std::unique(unique.begin(), unique.end()), unique.end());
    static_assert(sizeof(long long) == sizeof(int64_t), "period width");
    const int64_t *owned = n == ctx->d.ny ? periods + ctx->d.row0 : periods;
    if (dev_upload(ctx, ctx->flow_periods, owned, (size_t)ctx->d.rows * 8)) return 1;
    ctx->flow_unique.swap(unique);
    ctx->flow = true;
    return 0;
}

int fds_last_flow_shifts(fds_ctx *ctx, int64_t *shifts) {
    if (!ctx || !shifts) return fail(ctx, "fds_last_flow_shifts: null argument");
    *shifts = ctx->flow_shifts;
    return 0;
}

int fds_snapshot_async(fds_ctx *ctx, int32_t component, int32_t stride_x, int32_t stride_y,
                       int32_t slot) {
    NvtxRange nvtx_range("fds:snapshot");
    if (!ctx) return fail(ctx, "fds_snapshot_async: null context");
    if (component < 0 || component >= ctx->ncomp)
        return fail(ctx, "fds_snapshot_async: bad component");
    if (stride_x < 1 || stride_y < 1) return fail(ctx, "fds_snapshot_async: strides must be >= 1");
    if (slot < 0 || slot > 1) return fail(ctx, "fds_snapshot_async: slot must be 0 or 1");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    const long long fx = (ctx->d.nx + stride_x - 1) / stride_x;
    const long long fy = (ctx->d.rows + stride_y - 1) / stride_y;
    const size_t bytes = (size_t)(fx * fy) * 8;
    if (!ctx->ev_frame_ready[slot]) {
        FDS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_frame_ready[slot], cudaEventDisableTiming));
        FDS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_frame_done[slot], cudaEventDisableTiming));
    }
    // the slot may still be on its way to the host from an earlier snapshot
    FDS_CUDA(ctx, cudaEventSynchronize(ctx->ev_frame_done[slot]));
    if (ctx->frame_dev_cap[slot] < bytes) {
        if (ctx->frame_dev[slot]) {
            cudaFree(ctx->frame_dev[slot]);
            ctx->device_bytes -= (long long)ctx->frame_dev_cap[slot];
        }
        ctx->frame_dev[slot] = nullptr;
        ctx->frame_dev_cap[slot] = 0;
        if (dev_alloc(ctx, (void **)&ctx->frame_dev[slot], bytes, false)) return 1;
        ctx->frame_dev_cap[slot] = bytes;
    }
    if (ctx->frame_host_cap[slot] < bytes) {
        if (ctx->frame_host[slot]) cudaFreeHost(ctx->frame_host[slot]);
        ctx->frame_host[slot] = nullptr;
        ctx->frame_host_cap[slot] = 0;
        FDS_CUDA(ctx, cudaHostAlloc(&ctx->frame_host[slot], bytes, cudaHostAllocPortable));
        ctx->frame_host_cap[slot] = bytes;
    }
    SnapshotArgs a{};
    a.state = origin(ctx, ctx->cur, component);
    a.frame = ctx->frame_dev[slot];
    a.nx = ctx->d.nx;
    a.fx = fx;
    a.fy = fy;
    a.stride_x = stride_x;
    a.stride_y = stride_y;
    const long long blocks = std::min<long long>((fx * fy + 255) / 256, 148 * 16);
    snapshot_kernel<<<(unsigned)std::max<long long>(blocks, 1), 256, 0, ctx->stream>>>(a);
    FDS_CUDA(ctx, cudaGetLastError());
    // the copy to the host runs on the drain stream, behind whatever the caller enqueues next
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_frame_ready[slot], ctx->stream));
    FDS_CUDA(ctx, cudaStreamWaitEvent(ctx->drain, ctx->ev_frame_ready[slot], 0));
    FDS_CUDA(ctx, cudaMemcpyAsync(ctx->frame_host[slot], ctx->frame_dev[slot], bytes,
                                  cudaMemcpyDeviceToHost, ctx->drain));
    FDS_CUDA(ctx, cudaEventRecord(ctx->ev_frame_done[slot], ctx->drain));
    ctx->frame_n[slot] = fx * fy;
    return 0;
}

int fds_snapshot_wait(fds_ctx *ctx, int32_t slot, double *frame, int64_t n) {
    if (!ctx || !frame) return fail(ctx, "fds_snapshot_wait: null argument");
    if (slot < 0 || slot > 1 || !ctx->ev_frame_done[slot] || ctx->frame_n[slot] == 0)
        return fail(ctx, "fds_snapshot_wait: no snapshot pending in this slot");
    if (n != ctx->frame_n[slot])
        return fail(ctx, "fds_snapshot_wait: frame size differs from the snapshot taken");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    FDS_CUDA(ctx, cudaEventSynchronize(ctx->ev_frame_done[slot]));
    memcpy(frame, ctx->frame_host[slot], (size_t)n * 8);
    ctx->frame_n[slot] = 0;
    return 0;
}

int fds_comm_unique_id(uint8_t id[128]) {
    if (const char *e = load_nccl()) return fail(nullptr, std::string("fds_comm_unique_id: ") + e);
    nccl_unique_id uid;
    int rc = g_nccl.GetUniqueId(&uid);
    if (rc) return fail(nullptr, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    memcpy(id, uid.internal, 128);
    return 0;
}

int fds_comm_init(fds_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    if (!ctx || !id) return fail(ctx, "fds_comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "fds_comm_init: bad rank/world");
    if (world > 1 && ctx->d.halo_rows < 1)
        return fail(ctx, "fds_comm_init: a multi-slab run needs halo_rows >= 1");
    if (const char *e = load_nccl()) return fail(ctx, std::string("fds_comm_init: ") + e);
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    nccl_unique_id uid;
    memcpy(uid.internal, id, 128);
    FDS_NCCL(ctx, g_nccl.CommInitRank(&ctx->comm, world, uid, rank));
    ctx->rank = rank;
    ctx->world = world;
    return 0;
}

int fds_peer_export(fds_ctx *ctx, uint8_t *handles) {
    if (!ctx || !handles) return fail(ctx, "fds_peer_export: null argument");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int at = 0;
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 3; ++c, ++at) {
            cudaIpcMemHandle_t h;
            memset(&h, 0, sizeof(h));
            if (ctx->buf[b][c]) FDS_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->buf[b][c]));
            memcpy(handles + 64 * at, &h, 64);
        }
    cudaIpcMemHandle_t h;
    FDS_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->flags));
    memcpy(handles + 64 * at, &h, 64);
    return 0;
}

int fds_peer_import(fds_ctx *ctx, int32_t side, const uint8_t *handles, int64_t neighbour_rows) {
    if (!ctx || !handles) return fail(ctx, "fds_peer_import: null argument");
    if (side < 0 || side > 1) return fail(ctx, "fds_peer_import: side must be 0 (lower) or 1 (upper)");
    if (ctx->peer_open[side]) return fail(ctx, "fds_peer_import: side already imported");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    int at = 0;
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 3; ++c, ++at) {
            if (!ctx->buf[b][c]) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + 64 * at, 64);
            FDS_CUDA(ctx, cudaIpcOpenMemHandle(&ctx->peer_base[side][b][c], h,
                                               cudaIpcMemLazyEnablePeerAccess));
        }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * at, 64);
    void *flags = nullptr;
    FDS_CUDA(ctx, cudaIpcOpenMemHandle(&flags, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_flags[side] = (unsigned *)flags;
    ctx->peer_rows[side] = neighbour_rows;
    ctx->peer_open[side] = true;
    return 0;
}

// ---- coupled fields ---------------------------------------------------------------------------------

int fds_group_create(fds_ctx *const *members, int32_t n_members, fds_group **out) {
    if (!members || !out || n_members < 1) return gfail(nullptr, "fds_group_create: bad arguments");
    *out = nullptr;
    for (int m = 0; m < n_members; ++m) {
        if (!members[m]) return gfail(nullptr, "fds_group_create: null member");
        if (members[m]->dims != 1)
            return gfail(nullptr, "fds_group_create: device groups are built from 1-D fields");
        if (members[m]->d.device != members[0]->d.device || members[m]->d.nx != members[0]->d.nx)
            return gfail(nullptr, "fds_group_create: members must share the device and the line");
        if (members[m]->d.kernel == 1)
            return gfail(nullptr, "fds_group_create: members must use the default 1-D kernel");
    }
    fds_group *g = new fds_group();
    g->members.assign(members, members + n_members);
    g->n = members[0]->d.nx;
    *out = g;
    return 0;
}

void fds_group_destroy(fds_group *g) {
    if (!g) return;
    if (!g->members.empty()) {
        cudaSetDevice(g->members[0]->d.device);
        cudaStreamSynchronize(g->members[0]->stream);
    }
    if (g->d_tables) cudaFree(g->d_tables);
    for (auto &it : g->interactions) {
        if (it.acc) cudaFree(it.acc);
        if (it.aux) cudaFree(it.aux);
        if (it.law_buffers) cudaFree(it.law_buffers);
        if (it.law_max) cudaFree(it.law_max);
        if (it.law_count) cudaFree(it.law_count);
    }
    delete g;
}

const char *fds_group_last_error(const fds_group *g) {
    return g ? g->err.c_str() : g_create_error.c_str();
}

int fds_group_add_linear(fds_group *g, int32_t src_member, int32_t src_component,
                         int32_t dst_member, int32_t dst_component, double scale, int32_t additive,
                         int32_t accumulate, int64_t stepping, const double *accumulated) {
    if (!g) return gfail(g, "fds_group_add_linear: null group");
    if (!group_member_ok(g, src_member, src_component) ||
        !group_member_ok(g, dst_member, dst_component) || stepping < 1)
        return gfail(g, "fds_group_add_linear: bad member, component or stepping");
    FDS_GCUDA(g, cudaSetDevice(g->members[0]->d.device));
    fds_group::Interaction it;
    it.kind = 0;
    it.src_member = src_member; it.src_comp = src_component;
    it.dst_member = dst_member; it.dst_comp = dst_component;
    it.scale = scale;
    it.additive = additive ? 1 : 0;
    it.accumulate = accumulate ? 1 : 0;
    it.stepping = stepping;
    if (it.accumulate && group_alloc(g, (void **)&it.acc, sizeof(double) * (size_t)g->n, accumulated))
        return 1;
    g->interactions.push_back(it);
    return 0;
}

int fds_group_add_viscous_heating(fds_group *g, int32_t sound_member, int32_t heat_member,
                                  const double *density, const double *gradient_factor,
                                  const double *gain, double dt, int32_t accumulate,
                                  int64_t stepping, const double *accumulated) {
    if (!g || !density || !gradient_factor || !gain)
        return gfail(g, "fds_group_add_viscous_heating: null argument");
    if (!group_member_ok(g, sound_member, 1) || !group_member_ok(g, heat_member, 0) || stepping < 1)
        return gfail(g, "fds_group_add_viscous_heating: bad member or stepping");
    if (g->members[(size_t)sound_member]->thermal || !g->members[(size_t)heat_member]->thermal)
        return gfail(g, "fds_group_add_viscous_heating: needs an acoustic source and a thermal target");
    FDS_GCUDA(g, cudaSetDevice(g->members[0]->d.device));
    fds_group::Interaction it;
    it.kind = 1;
    it.src_member = sound_member; it.src_comp = 1;     // velocity
    it.dst_member = heat_member; it.dst_comp = 0;      // temperature
    it.additive = 1;
    it.accumulate = accumulate ? 1 : 0;
    it.stepping = stepping;
    it.dt = dt;
    const size_t n = (size_t)g->n;
    std::vector<double> packed(3 * n);
    memcpy(packed.data(), density, n * 8);
    memcpy(packed.data() + n, gradient_factor, n * 8);
    memcpy(packed.data() + 2 * n, gain, n * 8);
    if (group_alloc(g, (void **)&it.aux, sizeof(double) * 3 * n, packed.data())) return 1;
    if (it.accumulate && group_alloc(g, (void **)&it.acc, sizeof(double) * n, accumulated)) return 1;
    g->interactions.push_back(it);
    return 0;
}

int fds_group_add_material_law(fds_group *g, int32_t src_member, int32_t src_component,
                               int32_t dst_member, int32_t parameter, int32_t law, double p0,
                               double p1, int32_t has_threshold, double threshold,
                               int64_t stepping, const double *statics, const double *last,
                               const double *scalars) {
    if (!g || !statics || !scalars) return gfail(g, "fds_group_add_material_law: null argument");
    if (!group_member_ok(g, src_member, src_component) || !group_member_ok(g, dst_member, 0) ||
        parameter < 0 || parameter > 2 || stepping < 1 || (law != kLawExponential && law != kLawPower))
        return gfail(g, "fds_group_add_material_law: bad member, parameter, law or stepping");
    fds_ctx *target = g->members[(size_t)dst_member];
    if (!target->cell_tab)
        return gfail(g, "fds_group_add_material_law: the target field needs per-cell coefficients "
                        "(fds_upload_cell_table)");
    FDS_GCUDA(g, cudaSetDevice(g->members[0]->d.device));
    fds_group::Interaction it;
    it.kind = 2;
    it.src_member = src_member; it.src_comp = src_component;
    it.dst_member = dst_member;
    it.stepping = stepping;
    const size_t n = (size_t)g->n;
    std::vector<double> packed(5 * n, 0.0);
    memcpy(packed.data(), statics, 3 * n * 8);
    if (last) memcpy(packed.data() + 3 * n, last, n * 8);
    if (group_alloc(g, (void **)&it.law_buffers, sizeof(double) * 5 * n, packed.data())) return 1;
    if (group_alloc(g, (void **)&it.law_max, sizeof(unsigned long long), nullptr)) return 1;
    if (group_alloc(g, (void **)&it.law_count, sizeof(int), nullptr)) return 1;
    LawArgs &a = it.law;
    a.statics = it.law_buffers;
    a.last = it.law_buffers + 3 * n;
    a.factors = it.law_buffers + 4 * n;
    a.max_bits = it.law_max;
    a.reassemblies = it.law_count;
    a.n = g->n;
    a.law = law;
    a.p0 = p0;
    a.p1 = p1;
    a.p2 = 1.0 - p0;                // exponential: (1 - a), as Python evaluates it
    a.has_threshold = has_threshold ? 1 : 0;
    a.threshold = threshold;
    a.target_model = target->thermal ? kTargetThermal1D : kTargetAcoustic1D;
    a.parameter = parameter;
    a.k_dtdx = scalars[0];
    a.k_dtdx2 = scalars[1];
    a.k_inv_dx = scalars[2];
    g->interactions.push_back(it);
    return 0;
}

int fds_group_read(fds_group *g, int32_t interaction, double *values, int64_t *reassemblies) {
    if (!g || interaction < 0 || interaction >= (int)g->interactions.size())
        return gfail(g, "fds_group_read: bad interaction");
    FDS_GCUDA(g, cudaSetDevice(g->members[0]->d.device));
    FDS_GCUDA(g, cudaStreamSynchronize(g->members[0]->stream));
    const fds_group::Interaction &it = g->interactions[(size_t)interaction];
    const size_t bytes = sizeof(double) * (size_t)g->n;
    if (reassemblies) *reassemblies = 0;
    if (it.kind == 2) {
        if (values) FDS_GCUDA(g, cudaMemcpy(values, it.law.last, bytes, cudaMemcpyDeviceToHost));
        int count = 0;
        FDS_GCUDA(g, cudaMemcpy(&count, it.law_count, sizeof(int), cudaMemcpyDeviceToHost));
        if (reassemblies) *reassemblies = count;
    } else if (values) {
        if (it.acc) FDS_GCUDA(g, cudaMemcpy(values, it.acc, bytes, cudaMemcpyDeviceToHost));
        else memset(values, 0, bytes);
    }
    return 0;
}

int fds_group_step(fds_group *g, int64_t first_step, int64_t n_steps, double *const *probes_out) {
    if (!g) return gfail(g, "fds_group_step: null group");
    if (n_steps <= 0) return 0;
    NvtxRange nvtx_range("fds:group step");
    const size_t nm = g->members.size();
    FDS_GCUDA(g, cudaSetDevice(g->members[0]->d.device));
    cudaStream_t stream = g->members[0]->stream;
    std::vector<StepTables> tables(nm);
    std::vector<cudaStream_t> own(nm);
    long long chunk = n_steps;
    auto member_fail = [&](fds_ctx *ctx) { return gfail(g, "fds_group_step: " + ctx->err); };
    for (size_t m = 0; m < nm; ++m) {
        fds_ctx *ctx = g->members[m];
        if (!ctx->map_uploaded) return gfail(g, "fds_group_step: material map not uploaded");
        if (ctx->n_signals > 0 && (first_step < ctx->sig_first ||
                                   first_step + n_steps > ctx->sig_first + ctx->sig_steps))
            return gfail(g, "fds_group_step: step range outside the uploaded signal window");
        if (ctx->n_slots > 0 && (!probes_out || !probes_out[m]))
            return gfail(g, "fds_group_step: probes are configured but probes_out is NULL");
        if (refresh_flags(ctx) || ensure_ring(ctx, n_steps)) return member_fail(ctx);
        FDS_GCUDA(g, cudaStreamSynchronize(ctx->stream));   // everything runs on the first stream
        if (ctx->n_slots > 0) chunk = std::min(chunk, ctx->ring_half);
        ctx->last_launches = 0;
    }
    // the members' launches go to one stream: their order on it is the order of sim_step
    for (size_t m = 0; m < nm; ++m) {
        own[m] = g->members[m]->stream;
        g->members[m]->stream = stream;
        tables[m] = make_tables(g->members[m]);
    }
    int rc = 0;
    const bool pair = nm == 2 && !getenv("FDS_NO_PAIR_KERNEL");
    if (pair) {
        if (!g->d_tables && group_alloc(g, (void **)&g->d_tables, 2 * sizeof(StepTables), nullptr))
            rc = 1;
        if (!rc && cudaMemcpyAsync(g->d_tables, tables.data(), 2 * sizeof(StepTables),
                                   cudaMemcpyHostToDevice, stream) != cudaSuccess)
            rc = gfail(g, "fds_group_step: table upload failed");
        if (!rc) cudaStreamSynchronize(stream);      // `tables` is pageable host memory
    }
    cudaError_t e = cudaEventRecord(g->members[0]->ev_t0, stream);
    for (long long done = 0; done < n_steps && !rc && e == cudaSuccess; done += chunk) {
        const long long count = std::min(chunk, n_steps - done);
        for (long long s = 0; s < count && !rc; ++s) {
            const long long step = first_step + done + s;
            if (pair) rc = group_pair_step(g, step, s, stream);
            for (size_t m = 0; m < nm && !rc && !pair; ++m) {
                fds_ctx *ctx = g->members[m];
                rc = group_member_step(ctx, tables[m], step - ctx->sig_first, s);
                if (rc) g->err = "fds_group_step: " + ctx->err;
            }
            for (auto &it : g->interactions)
                if (!rc) rc = group_apply(g, it, step, stream);
        }
        for (size_t m = 0; m < nm && !rc; ++m) {
            fds_ctx *ctx = g->members[m];
            if (ctx->n_slots == 0) continue;
            e = cudaMemcpyAsync(probes_out[m] + done * ctx->n_slots, ctx->ring.ptr,
                                (size_t)count * ctx->n_slots * 8, cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) break;
        }
    }
    if (e == cudaSuccess) e = cudaEventRecord(g->members[0]->ev_t1, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    for (size_t m = 0; m < nm; ++m) g->members[m]->stream = own[m];
    g->members[0]->timed = true;
    if (rc) return 1;
    if (e != cudaSuccess) return gfail(g, std::string("fds_group_step: ") + cudaGetErrorString(e));
    return 0;
}

int fds_slab_init(fds_ctx *ctx, int32_t rank, int32_t world) {
    if (!ctx) return fail(ctx, "fds_slab_init: null context");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "fds_slab_init: bad rank/world");
    if (world > 1 && ctx->d.halo_rows < 1)
        return fail(ctx, "fds_slab_init: a multi-slab run needs halo_rows >= 1");
    if (ctx->comm) return fail(ctx, "fds_slab_init: the context already joined an NCCL communicator");
    ctx->rank = rank;
    ctx->world = world;
    ctx->local_peers = world > 1;
    return 0;
}

int fds_peer_connect(fds_ctx *ctx, int32_t side, fds_ctx *neighbour) {
    if (!ctx || !neighbour) return fail(ctx, "fds_peer_connect: null argument");
    if (side < 0 || side > 1) return fail(ctx, "fds_peer_connect: side must be 0 (lower) or 1 (upper)");
    if (!ctx->local_peers) return fail(ctx, "fds_peer_connect: call fds_slab_init first");
    if (ctx->peer_open[side]) return fail(ctx, "fds_peer_connect: side already connected");
    if (neighbour->d.nx != ctx->d.nx || neighbour->d.halo_rows != ctx->d.halo_rows ||
        neighbour->d.model != ctx->d.model)
        return fail(ctx, "fds_peer_connect: the neighbour is not a slab of the same field");
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    if (neighbour->d.device != ctx->d.device) {
        int can = 0;
        FDS_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->d.device, neighbour->d.device));
        if (!can) return fail(ctx, "fds_peer_connect: no peer access between the two devices");
        const cudaError_t e = cudaDeviceEnablePeerAccess(neighbour->d.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(ctx, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 3; ++c) ctx->peer_base[side][b][c] = neighbour->buf[b][c];
    ctx->peer_flags[side] = neighbour->flags;
    ctx->peer_rows[side] = neighbour->d.rows;
    ctx->local_neighbour[side] = neighbour;
    ctx->peer_open[side] = true;
    return 0;
}

int fds_last_step_ms(fds_ctx *ctx, double *ms) {
    if (!ctx || !ms) return fail(ctx, "fds_last_step_ms: null argument");
    if (!ctx->timed) return fail(ctx, "fds_last_step_ms: no step call recorded yet");
    float f = 0.f;
    FDS_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
    *ms = (double)f;
    return 0;
}

int fds_last_launch_info(fds_ctx *ctx, int64_t *launches, int64_t *steps_per_launch,
                         const char **kernel_name) {
    if (!ctx) return fail(ctx, "fds_last_launch_info: null context");
    if (launches) *launches = ctx->last_launches;
    if (steps_per_launch) *steps_per_launch = ctx->last_steps_per_launch;
    if (kernel_name) *kernel_name = ctx->last_kernel;
    return 0;
}

int64_t fds_device_bytes(const fds_ctx *ctx) { return ctx ? ctx->device_bytes : 0; }

int fds_stream_stats(fds_ctx *ctx, int64_t out[8]) {
    if (!ctx) return fail(ctx, "fds_stream_stats: null context");
    for (int k = 0; k < 8; ++k) out[k] = 0;
    if (!ctx->stream_stats) return 0;
    FDS_CUDA(ctx, cudaSetDevice(ctx->d.device));
    FDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FDS_CUDA(ctx, cudaMemcpy(out, ctx->stream_stats, sizeof(int64_t) * 8, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
