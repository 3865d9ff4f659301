// Row-band split of ONE very large frame over several GPUs (BASELINE.json configs[4]): the cost and
// cross-aggregation stages of CBLSM ComputeAD -> CrossAggregator::Aggregate (CBLSM/CBLSM.h:327, cross_aggregator.cpp:89-118).
//
// GPU g owns image rows [r0, r1) and keeps HALO = 34 rows (the maximum arm length, cross_aggregator.cpp:151 with the
// reference's L1) beyond each interior edge.  A horizontal pass is row-local and runs on the OWNED rows only; a vertical
// pass reads 34 rows beyond the band, i.e. the neighbours' current cost.  Those rows travel ONE-SIDED over NVLink:
//   * a band lives in one device allocation [header | volume 0 | volume 1]; neighbours map it (CUDA IPC across processes,
//     plain peer access inside one process), so a rank can write straight into its neighbours' halo rows;
//   * before a vertical pass each rank copies its 34 top / bottom owned rows of the pass's input into the lower halo of the
//     rank above / the upper halo of the rank below (copy engines, on a side stream) and then stores the exchange's
//     sequence number into that neighbour's header; the neighbour's vertical pass is queued behind a one-warp kernel
//     that spins on its own header until both neighbours' numbers have arrived.  No collective, no host synchronisation,
//     no NCCL in the data path;
//   * when the pass before the exchange is horizontal, its edge rows are computed first, the copies start, and the
//     interior rows are computed while they fly;
//   * nobody but the neighbours ever writes a halo row: vertical passes store owned rows only (st_lo / st_hi of the pass
//     kernels), so a push may land while the receiver is still busy with the previous pass.
// Ordering argument for the write-after-read hazard (a push overwrites halo rows the receiver read two sweeps ago): a
// rank can only push for exchange e after finishing the pass that needed exchange e-1 from the same neighbour, which
// that neighbour issued after finishing ITS pass before -- so the halo rows being overwritten were consumed at least
// one full exchange earlier.  Volumes 0 and 1 alternate every sweep on all ranks in lockstep.
#include "common.cuh"

namespace {

constexpr int BAND_HALO = 34;
constexpr size_t BAND_HEADER = 256;  // bytes reserved in front of the volumes

struct BandHeader {
    int from_upper, from_lower, error, pad;
};

struct PeerView {  // a neighbour's band as this rank sees it
    unsigned char* base = nullptr;
    int top = 0, bottom = 0, owned = 0;
    bool ipc = false;
};

struct HandleBody {  // what smt_band_export serialises (fits SMT_BAND_HANDLE_BYTES)
    cudaIpcMemHandle_t mem;
    int32_t cols, D, top, bottom, owned, device;
    uint32_t magic;
};
static_assert(sizeof(HandleBody) <= SMT_BAND_HANDLE_BYTES, "smt_band_handle too small");
constexpr uint32_t BAND_MAGIC = 0x534d5442u;

size_t band_vol_bytes(int nrows, int cols, int D) { return (((size_t)nrows * cols * D * sizeof(float)) + 255) & ~size_t(255); }

__global__ void band_signal_kernel(int* flag, int seq) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(flag) = seq;
    __threadfence_system();
}

// spins until both neighbours' sequence numbers have arrived (or ~4 s have passed: then the header's error word is set
// and the pass behind this kernel runs on stale halos -- smt_band_sync reports it)
__global__ void band_wait_kernel(BandHeader* h, int want_upper, int want_lower, int seq) {
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    volatile int* fu = &h->from_upper;
    volatile int* fl = &h->from_lower;
    while ((want_upper && *fu - seq < 0) || (want_lower && *fl - seq < 0)) {
        if (clock64() - t0 > 8000000000ll) {
            h->error = 1;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace

struct smt_band {
    smt_ctx* ctx = nullptr;
    int image_rows = 0, cols = 0, D = 0, rank = 0, world = 1;
    int r0 = 0, r1 = 0, top = 0, bottom = 0;
    unsigned char* base = nullptr;  // [header | vol 0 | vol 1]
    PeerView upper, lower;
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_pushed = nullptr;
    int seq = 0;
    double exchange_bytes = 0;  // pushed by this rank since creation
    int owned() const { return r1 - r0; }
    int nrows() const { return top + owned() + bottom; }
    size_t vol_bytes() const { return band_vol_bytes(nrows(), cols, D); }
    float* vol(int k) const { return reinterpret_cast<float*>(base + BAND_HEADER + (size_t)k * vol_bytes()); }
    BandHeader* header() const { return reinterpret_cast<BandHeader*>(base); }
    float* peer_vol(const PeerView& p, int k) const {
        return reinterpret_cast<float*>(p.base + BAND_HEADER + (size_t)k * band_vol_bytes(p.top + p.owned + p.bottom, cols, D));
    }
};

// copy my edge rows of volume k into the neighbours' halo rows of their volume k, then tell them (sequence number)
static smt_status_t band_push(smt_band* b, int k) {
    smt_ctx* ctx = b->ctx;
    if (!b->upper.base && !b->lower.base) return SMT_OK;
    const size_t row = (size_t)b->cols * b->D;  // floats per row
    const size_t bytes = (size_t)BAND_HALO * row * sizeof(float);
    b->seq += 1;
    SMT_CUDA(ctx, cudaEventRecord(b->ev_ready, ctx->stream));
    SMT_CUDA(ctx, cudaStreamWaitEvent(b->copy, b->ev_ready, 0));
    if (b->upper.base) {  // my first owned rows -> the lower halo of the rank above
        float* dst = b->peer_vol(b->upper, k) + (size_t)(b->upper.top + b->upper.owned) * row;
        SMT_CUDA(ctx, cudaMemcpyAsync(dst, b->vol(k) + (size_t)b->top * row, bytes, cudaMemcpyDeviceToDevice, b->copy));
        band_signal_kernel<<<1, 1, 0, b->copy>>>(&reinterpret_cast<BandHeader*>(b->upper.base)->from_lower, b->seq);
        b->exchange_bytes += (double)bytes;
    }
    if (b->lower.base) {  // my last owned rows -> the upper halo of the rank below
        float* dst = b->peer_vol(b->lower, k);
        SMT_CUDA(ctx, cudaMemcpyAsync(dst, b->vol(k) + (size_t)(b->top + b->owned() - BAND_HALO) * row, bytes, cudaMemcpyDeviceToDevice,
                                      b->copy));
        band_signal_kernel<<<1, 1, 0, b->copy>>>(&reinterpret_cast<BandHeader*>(b->lower.base)->from_upper, b->seq);
        b->exchange_bytes += (double)bytes;
    }
    ctx->launches += (b->upper.base ? 1 : 0) + (b->lower.base ? 1 : 0);
    SMT_CUDA(ctx, cudaGetLastError());
    SMT_CUDA(ctx, cudaEventRecord(b->ev_pushed, b->copy));
    return SMT_OK;
}

// queue the main stream behind this exchange: my own copies (their source rows are overwritten two sweeps later) and the
// neighbours' sequence numbers
static smt_status_t band_wait(smt_band* b) {
    smt_ctx* ctx = b->ctx;
    if (!b->upper.base && !b->lower.base) return SMT_OK;
    SMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, b->ev_pushed, 0));
    SMT_LAUNCH(ctx, band_wait_kernel, 1, 32, 0, b->header(), b->upper.base ? 1 : 0, b->lower.base ? 1 : 0, b->seq);
    return SMT_OK;
}

extern "C" {

smt_status_t smt_band_create(smt_ctx* ctx, int32_t image_rows, int32_t cols, int32_t D, int32_t rank, int32_t world, smt_band** out) {
    if (!ctx || !out) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, image_rows > 0 && cols > 0 && D > 0 && world >= 1 && rank >= 0 && rank < world, SMT_ERR_INVALID_ARG,
                "smt_band_create: bad argument");
    // the split is a pure function of (rows, world): every rank sees the same bands and fails the same way
    const int per = (image_rows + world - 1) / world;
    for (int r = 0; r < world; ++r) {
        const int a = r * per < image_rows ? r * per : image_rows, e = (r + 1) * per < image_rows ? (r + 1) * per : image_rows;
        if (world > 1 && e - a < BAND_HALO)
            return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "smt_band_create: band %d of %d would own %d rows, fewer than the %d-row halo; use fewer GPUs",
                            r, world, e - a, BAND_HALO);
    }
    DeviceGuard dev_guard(ctx->device);
    smt_band* b = new smt_band();
    b->ctx = ctx;
    b->image_rows = image_rows;
    b->cols = cols;
    b->D = D;
    b->rank = rank;
    b->world = world;
    b->r0 = rank * per < image_rows ? rank * per : image_rows;
    b->r1 = (rank + 1) * per < image_rows ? (rank + 1) * per : image_rows;
    b->top = rank > 0 ? BAND_HALO : 0;
    b->bottom = rank < world - 1 ? BAND_HALO : 0;
    const size_t total = BAND_HEADER + 2 * b->vol_bytes();
    if (cudaMalloc(&b->base, total) != cudaSuccess) {
        cudaGetLastError();
        delete b;
        return smt_fail(ctx, SMT_ERR_OOM, "smt_band_create: device allocation of %zu bytes", total);
    }
    if (cudaMemsetAsync(b->base, 0, BAND_HEADER, ctx->stream) != cudaSuccess ||
        cudaStreamCreateWithFlags(&b->copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_pushed, cudaEventDisableTiming) != cudaSuccess) {
        smt_band_destroy(b);
        return smt_fail(ctx, SMT_ERR_CUDA, "smt_band_create: stream / event creation failed");
    }
    cudaStreamSynchronize(ctx->stream);  // the header is zero before anyone can map it
    *out = b;
    return SMT_OK;
}

void smt_band_destroy(smt_band* b) {
    if (!b) return;
    DeviceGuard dev_guard(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->copy) cudaStreamSynchronize(b->copy);
    for (PeerView* p : {&b->upper, &b->lower})
        if (p->base && p->ipc) cudaIpcCloseMemHandle(p->base);
    if (b->ev_ready) cudaEventDestroy(b->ev_ready);
    if (b->ev_pushed) cudaEventDestroy(b->ev_pushed);
    if (b->copy) cudaStreamDestroy(b->copy);
    if (b->base) cudaFree(b->base);
    delete b;
}

smt_status_t smt_band_geometry(const smt_band* b, int32_t* r0, int32_t* r1, int32_t* top, int32_t* bottom) {
    if (!b) return SMT_ERR_INVALID_ARG;
    if (r0) *r0 = b->r0;
    if (r1) *r1 = b->r1;
    if (top) *top = b->top;
    if (bottom) *bottom = b->bottom;
    return SMT_OK;
}

smt_status_t smt_band_volume(smt_band* b, float** vol) {
    if (!b || !vol) return SMT_ERR_INVALID_ARG;
    *vol = b->vol(0);
    return SMT_OK;
}

smt_status_t smt_band_export(smt_band* b, smt_band_handle* out) {
    if (!b || !out) return SMT_ERR_INVALID_ARG;
    smt_ctx* ctx = b->ctx;
    DeviceGuard dev_guard(ctx->device);
    HandleBody h;
    memset(&h, 0, sizeof h);
    SMT_CUDA(ctx, cudaIpcGetMemHandle(&h.mem, b->base));
    h.cols = b->cols;
    h.D = b->D;
    h.top = b->top;
    h.bottom = b->bottom;
    h.owned = b->owned();
    h.device = ctx->device;
    h.magic = BAND_MAGIC;
    memset(out->bytes, 0, sizeof out->bytes);
    memcpy(out->bytes, &h, sizeof h);
    return SMT_OK;
}

static smt_status_t band_open(smt_band* b, const smt_band_handle* hnd, PeerView* view, bool is_upper) {
    smt_ctx* ctx = b->ctx;
    HandleBody h;
    memcpy(&h, hnd->bytes, sizeof h);
    if (h.magic != BAND_MAGIC || h.cols != b->cols || h.D != b->D || (is_upper ? h.bottom : h.top) != BAND_HALO || h.owned < BAND_HALO)
        return smt_fail(ctx, SMT_ERR_INVALID_ARG, "smt_band_connect: handle does not describe a matching %s neighbour", is_upper ? "upper" : "lower");
    void* p = nullptr;
    SMT_CUDA(ctx, cudaIpcOpenMemHandle(&p, h.mem, cudaIpcMemLazyEnablePeerAccess));
    view->base = static_cast<unsigned char*>(p);
    view->top = h.top;
    view->bottom = h.bottom;
    view->owned = h.owned;
    view->ipc = true;
    return SMT_OK;
}

smt_status_t smt_band_connect(smt_band* b, const smt_band_handle* upper, const smt_band_handle* lower) {
    if (!b) return SMT_ERR_INVALID_ARG;
    smt_ctx* ctx = b->ctx;
    SMT_REQUIRE(ctx, (upper != nullptr) == (b->rank > 0) && (lower != nullptr) == (b->rank < b->world - 1), SMT_ERR_INVALID_ARG,
                "smt_band_connect: pass exactly the handles of the existing neighbours (upper for rank > 0, lower for rank < world-1)");
    DeviceGuard dev_guard(ctx->device);
    if (upper) SMT_TRY(band_open(b, upper, &b->upper, true));
    if (lower) SMT_TRY(band_open(b, lower, &b->lower, false));
    return SMT_OK;
}

smt_status_t smt_band_connect_local(smt_band* b, smt_band* upper, smt_band* lower) {
    if (!b) return SMT_ERR_INVALID_ARG;
    smt_ctx* ctx = b->ctx;
    SMT_REQUIRE(ctx, (upper != nullptr) == (b->rank > 0) && (lower != nullptr) == (b->rank < b->world - 1), SMT_ERR_INVALID_ARG,
                "smt_band_connect_local: pass exactly the existing neighbours");
    DeviceGuard dev_guard(ctx->device);
    for (int side = 0; side < 2; ++side) {
        smt_band* n = side == 0 ? upper : lower;
        if (!n) continue;
        SMT_REQUIRE(ctx, n->cols == b->cols && n->D == b->D && n->world == b->world && n->rank == b->rank + (side == 0 ? -1 : 1),
                    SMT_ERR_INVALID_ARG, "smt_band_connect_local: not the neighbouring band of the same frame");
        if (n->ctx->device != ctx->device) {
            int can = 0;
            SMT_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, n->ctx->device));
            SMT_REQUIRE(ctx, can, SMT_ERR_UNSUPPORTED, "smt_band_connect_local: no peer access between the two devices");
            const cudaError_t e = cudaDeviceEnablePeerAccess(n->ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SMT_CUDA(ctx, e);
            cudaGetLastError();
        }
        PeerView* v = side == 0 ? &b->upper : &b->lower;
        v->base = n->base;
        v->top = n->top;
        v->bottom = n->bottom;
        v->owned = n->owned();
        v->ipc = false;
    }
    return SMT_OK;
}

// refresh the halo rows of volume 0 from the neighbours (and give them mine): what a caller needs before running its own
// vertical operator on a band; smt_band_aggregate_cross issues the same exchange before each of its vertical passes
smt_status_t smt_halo_exchange(smt_band* b) {
    if (!b) return SMT_ERR_INVALID_ARG;
    DeviceGuard dev_guard(b->ctx->device);
    SMT_TRY(band_push(b, 0));
    return band_wait(b);
}

smt_status_t smt_band_aggregate_cross(smt_band* b, const uint8_t* arms, const uint16_t* sup_hfirst, const uint16_t* sup_vfirst,
                                      int32_t iters, float vmax, smt_sum_mode_t mode) {
    if (!b) return SMT_ERR_INVALID_ARG;
    smt_ctx* ctx = b->ctx;
    SMT_REQUIRE(ctx, arms && sup_hfirst && sup_vfirst && iters >= 0, SMT_ERR_INVALID_ARG, "smt_band_aggregate_cross: bad argument");
    SMT_REQUIRE(ctx, mode == SMT_SUM_REFERENCE_ORDER || mode == SMT_SUM_FAST, SMT_ERR_INVALID_ARG,
                "smt_band_aggregate_cross: unknown sum mode");
    SMT_REQUIRE(ctx, (b->upper.base != nullptr) == (b->rank > 0) && (b->lower.base != nullptr) == (b->rank < b->world - 1),
                SMT_ERR_INVALID_ARG, "smt_band_aggregate_cross: connect the band to its neighbours first");
    DeviceGuard dev_guard(ctx->device);
    const int W = b->cols, D = b->D, top = b->top, own = b->owned(), nrows = b->nrows();
    if (mode == SMT_SUM_FAST) {
        int longest = 0;
        SMT_TRY(k_arms_max(ctx, arms, (size_t)nrows * W, &longest));
        if (longest > BAND_HALO)
            return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "smt_band_aggregate_cross: arms <= %d supported (halo width), the maps hold %d", BAND_HALO, longest);
    }
    const size_t row = (size_t)W * D;
    const size_t ntasks = (size_t)(nrows > W ? nrows : W) * ((D + 31) / 32);
    const bool neighbours = b->upper.base || b->lower.base;
    bool first = true;
    // rows [lo, hi) of the band buffers, one horizontal pass
    auto hpass = [&](int s, int k, int lo, int hi) -> smt_status_t {
        if (hi <= lo) return SMT_OK;
        const smt_status_t st = k_cross_sweep(ctx, iters, s, b->vol(k) + (size_t)lo * row, b->vol(k ^ 1) + (size_t)lo * row, arms + (size_t)lo * W * 4,
                                              sup_hfirst + (size_t)lo * W, sup_vfirst + (size_t)lo * W, hi - lo, W, D, vmax, mode, first, 0, 0, ntasks);
        first = false;
        return st;
    };
    int k = 0;  // volume holding the current cost
    for (int s = 0; s < 2 * iters; ++s) {
        const bool hfirst = (s / 2) % 2 == 0, second = s % 2 == 1;
        const bool horiz = hfirst != second;
        const bool next_vertical = s + 1 < 2 * iters && !((((s + 1) / 2) % 2 == 0) != ((s + 1) % 2 == 1));
        if (horiz) {
            if (next_vertical && neighbours && own >= 3 * BAND_HALO) {
                // edge rows first, so that their copies overlap the interior rows
                const int e0 = b->upper.base ? BAND_HALO : 0, e1 = b->lower.base ? BAND_HALO : 0;
                SMT_TRY(hpass(s, k, top, top + e0));
                SMT_TRY(hpass(s, k, top + own - e1, top + own));
                SMT_TRY(band_push(b, k ^ 1));
                SMT_TRY(hpass(s, k, top + e0, top + own - e1));
            } else {
                SMT_TRY(hpass(s, k, top, top + own));
                if (next_vertical && neighbours) SMT_TRY(band_push(b, k ^ 1));
            }
        } else {
            SMT_TRY(band_wait(b));
            SMT_TRY(k_cross_sweep(ctx, iters, s, b->vol(k), b->vol(k ^ 1), arms, sup_hfirst, sup_vfirst, nrows, W, D, vmax, mode, first, top,
                                  top + own, ntasks));
            first = false;
            if (next_vertical && neighbours) SMT_TRY(band_push(b, k ^ 1));
        }
        k ^= 1;
    }
    // 2*iters sweeps: the result is back in volume 0
    return SMT_OK;
}

smt_status_t smt_band_sync(smt_band* b, double* pushed_bytes) {
    if (!b) return SMT_ERR_INVALID_ARG;
    smt_ctx* ctx = b->ctx;
    DeviceGuard dev_guard(ctx->device);
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(b->copy));
    int err = 0;
    SMT_CUDA(ctx, cudaMemcpy(&err, &b->header()->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (pushed_bytes) *pushed_bytes = b->exchange_bytes;
    if (err) return smt_fail(ctx, SMT_ERR_CUDA, "smt_band_sync: a halo exchange timed out waiting for a neighbour (rank %d of %d)", b->rank, b->world);
    return SMT_OK;
}

}  // extern "C"
