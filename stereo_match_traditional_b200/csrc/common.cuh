// Internal header of libsmt_b200.so: context, device workspace, launch/error macros.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smt_b200.h"

enum SmtStage {
    ST_CENSUS = 0,
    ST_COST,
    ST_ARMS,
    ST_AGGREGATE,
    ST_SCANLINE,
    ST_WTA,
    ST_POST,
    ST_COPY,
    ST_COUNT
};

struct smt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // side stream of the pipeline entry points: image-sized, latency-bound kernels (arm construction) run on it beside
    // the volume-sized kernels of the main stream; created on first use
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_side[2] = {nullptr, nullptr};
    std::string last_error;
    uint64_t launches = 0;
    int sm_count = 148;
    size_t l2_bytes = 0;
    // grow-only cache of device blocks handed out by DevBuf
    struct Block {
        void* p;
        size_t bytes;
        bool used;
    };
    std::vector<Block> pool;
    // stage timing
    bool timing = false;
    double stage_ms[ST_COUNT] = {0};
    struct Pending {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Pending> pending;
    // exp LUT cache (host libm values, AD-Census.h:287-288)
    float lut_sigmaC = -1.f, lut_sigmaS = -1.f;
    float* d_lut = nullptr;  // 256 AD entries + 64 census entries
    int* d_flags = nullptr;  // small device scratch ints (arm-length reduction)
    int* d_redo = nullptr;   // redo state of the SMT_SUM_FAST aggregation (FX_ST_*), d_redo_cap list entries
    size_t d_redo_cap = 0;
    // pinned staging for small host<->device scalars
    int32_t* h_pinned = nullptr;
};

smt_status_t smt_fail(smt_ctx* ctx, smt_status_t st, const char* fmt, ...);
void* smt_pool_get(smt_ctx* ctx, size_t bytes);
void smt_pool_put(smt_ctx* ctx, void* p);

// Makes `device` current for the scope and restores the caller's device afterwards, so that an entry point never
// changes the host framework's current device behind its back.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Runs the enclosed launches on the context's side stream: fork() makes the side stream wait for everything queued on
// the main stream so far; the destructor switches back; join(k) makes the main stream wait for side-stream event k.
struct SideStream {
    smt_ctx* ctx;
    cudaStream_t main;
    bool ok = false;
    explicit SideStream(smt_ctx* c);
    ~SideStream() { ctx->stream = main; }
    void mark(int k) { if (ok) cudaEventRecord(ctx->ev_side[k], ctx->side); }
    static void join(smt_ctx* c, int k) { if (c->side) cudaStreamWaitEvent(c->stream, c->ev_side[k], 0); }
};

// RAII device buffer from the context's pool.
struct DevBuf {
    smt_ctx* ctx;
    void* p = nullptr;
    DevBuf(smt_ctx* c, size_t bytes) : ctx(c) { p = bytes ? smt_pool_get(c, bytes) : nullptr; }
    ~DevBuf() {
        if (p) smt_pool_put(ctx, p);
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
    bool ok() const { return p != nullptr; }
};

// Scoped stage timer (no-op unless smt_stage_timing_enable).
struct StageTimer {
    smt_ctx* ctx;
    int idx = -1;
    StageTimer(smt_ctx* c, int stage);
    ~StageTimer();
};

#define SMT_CUDA(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return smt_fail((ctx), e__ == cudaErrorMemoryAllocation ? SMT_ERR_OOM : SMT_ERR_CUDA,       \
                            "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));      \
    } while (0)

#define SMT_TRY(expr)                      \
    do {                                   \
        smt_status_t st__ = (expr);        \
        if (st__ != SMT_OK) return st__;   \
    } while (0)

#define SMT_LAUNCH(ctx, kernel, grid, block, smem, ...)                                                 \
    do {                                                                                                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                \
        (ctx)->launches++;                                                                              \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess)                                                                         \
            return smt_fail((ctx), SMT_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__, __LINE__, #kernel,  \
                            cudaGetErrorString(e__));                                                   \
    } while (0)

#define SMT_REQUIRE(ctx, cond, st, msg)                         \
    do {                                                        \
        if (!(cond)) return smt_fail((ctx), (st), "%s", (msg)); \
    } while (0)

#define SMT_ALLOC(ctx, name, bytes)                                                                  \
    DevBuf name((ctx), (bytes));                                                                     \
    if (((size_t)(bytes)) != 0 && !name.ok())                                                        \
    return smt_fail((ctx), SMT_ERR_OOM, "device allocation of %zu bytes", (size_t)(bytes))

// Redo state of a SMT_SUM_FAST aggregation, device ints shared by the fixed-point kernels (cross_agg_fx.cu, writers) and
// the float-pair passes queued behind them (cross_agg_seq.cu, readers): [ESCALATED] the float-pair passes take whole
// sweeps from now on, [COUNT] (line, disparity chunk) tasks on this sweep's redo list, [LIMIT] list length beyond which
// ESCALATED is raised, [TOTAL] tasks listed over all sweeps of the aggregation, [LIST..] the list.
enum { FX_ST_ESCALATED = 0, FX_ST_COUNT = 1, FX_ST_LIMIT = 2, FX_ST_TOTAL = 3, FX_ST_LIST = 4 };
// the context's redo state, large enough for `ntasks` list entries (grow-only; contents are lost when it grows)
int* smt_redo_state(smt_ctx* ctx, size_t ntasks);

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Host<->device marshalling for the SMT_HOST memspace: In() gives a device pointer holding a copy of a host
// array (or the pointer itself for SMT_DEVICE); Out() gives a device destination and copies back in finish().
struct Marshal {
    smt_ctx* ctx;
    smt_memspace_t space;
    struct Item {
        void* dev;
        void* host;
        size_t bytes;
        bool copy_back;
    };
    std::vector<Item> items;
    bool failed = false;
    Marshal(smt_ctx* c, smt_memspace_t s) : ctx(c), space(s) {}
    ~Marshal() {
        for (auto& it : items) smt_pool_put(ctx, it.dev);
    }
    template <typename T>
    const T* in(const T* p, size_t count) {
        if (!p || space == SMT_DEVICE) return p;
        void* d = smt_pool_get(ctx, count * sizeof(T));
        if (!d) {
            failed = true;
            return nullptr;
        }
        items.push_back({d, const_cast<T*>(p), count * sizeof(T), false});
        if (cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) failed = true;
        return static_cast<const T*>(d);
    }
    template <typename T>
    T* out(T* p, size_t count, bool also_in = false) {
        if (!p || space == SMT_DEVICE) return p;
        void* d = smt_pool_get(ctx, count * sizeof(T));
        if (!d) {
            failed = true;
            return nullptr;
        }
        items.push_back({d, p, count * sizeof(T), true});
        if (also_in &&
            cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
            failed = true;
        return static_cast<T*>(d);
    }
    smt_status_t check() {
        if (failed) return smt_fail(ctx, SMT_ERR_OOM, "host->device staging failed (allocation or copy)");
        return SMT_OK;
    }
    // copy outputs back and synchronise (host space only)
    smt_status_t finish() {
        if (space == SMT_DEVICE) return SMT_OK;
        for (auto& it : items)
            if (it.copy_back) SMT_CUDA(ctx, cudaMemcpyAsync(it.host, it.dev, it.bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SMT_OK;
    }
};

// ---- internal device-pointer entry points shared between translation units (all enqueue on ctx->stream) ----
smt_status_t k_f32_to_u8(smt_ctx* ctx, const float* in, size_t n, uint8_t* out);  // validates 0..255 integers
smt_status_t k_asw_crosscheck(smt_ctx* ctx, const float* dl, const float* dr, int H, int W, uint8_t* out);  // window_match.cu
smt_status_t k_u8_to_f32(smt_ctx* ctx, const uint8_t* in, size_t n, float* out);
smt_status_t k_bgr_to_gray(smt_ctx* ctx, const uint8_t* bgr, size_t n, uint8_t* gray);
smt_status_t k_ensure_lut(smt_ctx* ctx, float sigmaC, float sigmaS);
struct CensusMaps {
    // all uint64; see adcensus.cu for the definitions
    uint64_t* CL;   // [H][W]    left  own census (masked)
    uint64_t* CR;   // [H][W]    right own census (masked)
    uint64_t* CRx;  // [H][W+3]  right image census at virtual column x' in [-3, W-1]   (left-view matching)
    uint64_t* CLx;  // [H][W+4]  left  image census at virtual column x' in [0, W+3]    (right-view matching)
    uint64_t* M;    // [H][W]    validity mask of the anchor pixel
};
size_t census_maps_bytes(int H, int W);
void census_maps_carve(void* base, int H, int W, CensusMaps* m);
smt_status_t k_census_maps(smt_ctx* ctx, const uint8_t* L, const uint8_t* R, int H, int W, const CensusMaps& m);
smt_status_t k_adcensus_cost(smt_ctx* ctx, const uint8_t* L, const uint8_t* R, const CensusMaps& m, int H, int W, int D,
                             int view, float* cost, float* ad, float* census);
smt_status_t k_wta(smt_ctx* ctx, const float* vol, int H, int W, int D, float* disp);
smt_status_t k_cross_arms(smt_ctx* ctx, const uint8_t* bgr, int H, int W, int L1, int L2, int t1, int t2, uint8_t* arms,
                          uint16_t* sup0, uint16_t* sup1, uint32_t* scratch = nullptr);
// aggregates `vol` (modified: ping-pongs between vol and tmp); *result points at whichever buffer holds the answer
smt_status_t k_aggregate_cross(smt_ctx* ctx, float* vol, float* tmp, const uint8_t* arms, const uint16_t* sup0,
                               const uint16_t* sup1, int H, int W, int D, int iters, smt_sum_mode_t mode, float vmax, float** result,
                               float* wta_disp = nullptr);
// ---- cross aggregation building blocks (cross_agg*.cu).  st_lo, st_hi: only outputs at positions [st_lo, st_hi) of each
// walked line are stored (st_hi <= 0: the whole line); redo: the FX_ST_* state of the fixed-point sweep queued in front
bool fx_quantum(float vmax, float* inv_q, float* q);
smt_status_t k_fx_state_reset(smt_ctx* ctx, int* state, bool new_aggregation, size_t ntasks);
smt_status_t k_agg_fx_pass(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                           int horiz, int tmaps, float inv_q, float q, int* state, int st_lo = 0, int st_hi = 0);
smt_status_t k_agg_fx_pair(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                           int horiz, int tmaps, float inv_q, float q, int* state);
smt_status_t k_agg_seq_pass(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                            int horiz, bool nonneg, float bias, int tmaps, const int* redo, int st_lo = 0, int st_hi = 0);
smt_status_t k_agg_exact_pass(smt_ctx* ctx, const float* src, float* dst, const uint8_t* arms8, const uint16_t* s, int H, int W, int D,
                              int horiz, int st_lo = 0, int st_hi = 0);
smt_status_t k_transpose_maps(smt_ctx* ctx, const uint8_t* arms8, const uint16_t* sup, int H, int W, uint8_t* armsT, uint16_t* supT);
smt_status_t k_arms_max(smt_ctx* ctx, const uint8_t* arms8, size_t npix, int* longest);  // synchronises
// one of the 2*iters one-dimensional passes of CrossAggregator::Aggregate on a volume of `rows` rows (the per-sweep ABI and
// the row-band split are built on it); first_of_aggregation resets the redo state of the fast mode; ntasks_whole: (line,
// chunk) tasks of the largest pass of the aggregation when this call only covers part of a volume (sizes the redo list)
smt_status_t k_cross_sweep(smt_ctx* ctx, int iters, int sweep, const float* vol_in, float* vol_out, const uint8_t* arms,
                           const uint16_t* sup_hfirst, const uint16_t* sup_vfirst, int rows, int cols, int D, float vmax,
                           smt_sum_mode_t mode, bool first_of_aggregation, int st_lo = 0, int st_hi = 0, size_t ntasks_whole = 0);
smt_status_t k_scanline(smt_ctx* ctx, const float* cost, const float* gray, int H, int W, int D, int p1, int p2,
                        float* left, float* right, float* up, float* down, float* sum, float* disp, float* scratch_a,
                        float* scratch_b);
smt_status_t k_lrcheck(smt_ctx* ctx, float* dispL, const float* dispR, int W, int H, float gate, bool variant_and,
                       float* last, int32_t* occ, int32_t* mis, int32_t* counts_dev /*2 ints*/, uint8_t* cls_out = nullptr);
smt_status_t k_class_map_to_lists(smt_ctx* ctx, const unsigned char* cls, int W, int H, int32_t* occ, int32_t* mis, int32_t* counts_dev);
// multi-step refinement (refine.cu): cls = 0 valid / 1 occlusion / 2 mismatch
smt_status_t k_refine_region_voting(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* arms, int H, int W, int D, int ts, float th,
                                    int iters);
smt_status_t k_refine_interpolate(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* bgr, int H, int W, int max_len);
smt_status_t k_refine_discontinuity(smt_ctx* ctx, float* disp, const float* cost, int H, int W, int D, float thres);
smt_status_t k_remove_speckles(smt_ctx* ctx, float* disp, int W, int H, int diff, uint32_t min_area, int invalid_val);
smt_status_t k_median(smt_ctx* ctx, const float* in, float* out, int W, int H, int k);
