// The ASW driver's display chain on 8-bit maps (ASW/ASWeight.cpp:66-78):
//     CrossCheckDiaparity -> normalize(0, 255, NORM_MINMAX) -> filterSpeckles(0, 40, 2) -> medianBlur 5 -> FillImageNew
//     -> medianBlur 3
// normalize, filterSpeckles and medianBlur are OpenCV functions whose sources are not in the reference tree: they follow
// OpenCV's published behaviour and are pinned by golden vectors made with cv2 4.13 (oracle/make_golden_display.py);
// FillImageNew (ASW/ASW.h:434-511) follows the reference as written and is pinned against the compiled reference.
#include "common.cuh"
#include "unionfind.cuh"

// ---------------------------------------------------------------------------------------------- normalize, CV_8U -> CV_8U
__global__ void __launch_bounds__(256) minmax_u8_kernel(const uint8_t* __restrict__ v, size_t n, unsigned* __restrict__ out) {
    unsigned lo = 255u, hi = 0u;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const unsigned x = v[k];
        lo = min(lo, x);
        hi = max(hi, x);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}
// cv::normalize: scale = 255 * (1 / (max - min)) in double, 0 for a constant image; shift = -min * scale; convertTo
// evaluates src * (float)scale + (float)shift with one rounding and saturate_cast<uchar> rounds half to even
__global__ void __launch_bounds__(256) normalize_u8u8_kernel(const uint8_t* __restrict__ v, size_t n, const unsigned* __restrict__ mm,
                                                             uint8_t* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double lo = (double)mm[0], hi = (double)mm[1];
    const double scale = 255.0 * ((hi - lo) > 2.220446049250313e-16 ? 1.0 / (hi - lo) : 0.0);
    const double shift = 0.0 - lo * scale;
    int q = __float2int_rn(__fmaf_rn((float)v[p], (float)scale, (float)shift));
    out[p] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

static smt_status_t k_normalize_u8u8(smt_ctx* ctx, const uint8_t* src, size_t n, uint8_t* out) {
    SMT_ALLOC(ctx, mm, 2 * sizeof(unsigned));
    const unsigned init[2] = {255u, 0u};
    SMT_CUDA(ctx, cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    SMT_LAUNCH(ctx, minmax_u8_kernel, ctx->sm_count * 4, 256, 0, src, n, mm.as<unsigned>());
    SMT_LAUNCH(ctx, normalize_u8u8_kernel, cdiv((long long)n, 256), 256, 0, src, n, mm.as<unsigned>(), out);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- filterSpeckles, CV_8UC1
// OpenCV's flood fill finds the connected components of {4-neighbours, both != newVal, |a - b| <= maxDiff}; membership
// does not depend on the visiting order, so the lock-free union-find of the speckle remover (postproc.cu) gives the same
// components.  A component of at most maxSpeckleSize pixels becomes newVal.
__device__ __forceinline__ bool fs_conn(int a, int b, int new_val, int diff) { return a != new_val && b != new_val && abs(a - b) <= diff; }

// one block per row: label = start index of the pixel's horizontal run (pointer jumping in shared memory), -1 for newVal
__global__ void __launch_bounds__(256) fs_rows_kernel(const uint8_t* __restrict__ img, int W, int new_val, int diff, int* __restrict__ label,
                                                      int* __restrict__ size) {
    extern __shared__ int s_l[];
    const int i = blockIdx.x;
    const uint8_t* row = img + (size_t)i * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        const int v = row[j];
        int l = j;
        if (v == new_val) l = -1;
        else if (j > 0 && fs_conn(v, row[j - 1], new_val, diff)) l = j - 1;
        s_l[j] = l;
        size[(size_t)i * W + j] = 0;
    }
    __syncthreads();
    for (int span = 1; span < W; span <<= 1) {
        for (int j = threadIdx.x; j < W; j += blockDim.x) {
            const int l = ((volatile int*)s_l)[j];
            if (l >= 0) {
                const int ll = ((volatile int*)s_l)[l];
                if (ll != l) ((volatile int*)s_l)[j] = ll;
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        const int l = s_l[j];
        label[(size_t)i * W + j] = l < 0 ? -1 : i * W + l;
    }
}
// link the runs of adjacent rows; an edge that the pixel to the left already implies is skipped
__global__ void __launch_bounds__(256) fs_merge_kernel(const uint8_t* __restrict__ img, int W, int H, int new_val, int diff,
                                                       int* __restrict__ label) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
    if (i + 1 >= H) return;
    const int v = img[p], s = img[p + W];
    if (!fs_conn(v, s, new_val, diff)) return;
    if (j > 0) {
        const int vl = img[p - 1], sl = img[p + W - 1];
        if (fs_conn(v, vl, new_val, diff) && fs_conn(vl, sl, new_val, diff) && fs_conn(sl, s, new_val, diff)) return;
    }
    uf_union(label, (int)p, (int)(p + W));
}
__global__ void __launch_bounds__(256) fs_apply_kernel(const int* __restrict__ label, const int* __restrict__ size, size_t n, int max_size,
                                                       int new_val, uint8_t* __restrict__ img) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l = label[p];
    if (l >= 0 && size[l] <= max_size) img[p] = (uint8_t)new_val;
}

static smt_status_t k_filter_speckles_u8(smt_ctx* ctx, uint8_t* img, int H, int W, int new_val, int max_size, int max_diff) {
    const size_t n = (size_t)W * H;
    const unsigned nb = cdiv((long long)n, 256);
    SMT_ALLOC(ctx, label, n * sizeof(int));
    SMT_ALLOC(ctx, size, n * sizeof(int));
    if ((size_t)W * sizeof(int) > 200 * 1024) return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "filter_speckles: width %d too large", W);
    if ((size_t)W * sizeof(int) > 48 * 1024)
        SMT_CUDA(ctx, cudaFuncSetAttribute(fs_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(W * sizeof(int))));
    SMT_LAUNCH(ctx, fs_rows_kernel, H, 256, (size_t)W * sizeof(int), img, W, new_val, max_diff, label.as<int>(), size.as<int>());
    SMT_LAUNCH(ctx, fs_merge_kernel, nb, 256, 0, img, W, H, new_val, max_diff, label.as<int>());
    SMT_LAUNCH(ctx, uf_count_kernel, nb, 256, 0, label.as<int>(), n, size.as<int>());
    SMT_LAUNCH(ctx, fs_apply_kernel, nb, 256, 0, label.as<int>(), size.as<int>(), n, max_size, new_val, img);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- medianBlur, CV_8UC1
// median of the K x K window with BORDER_REPLICATE: the (K*K/2)-th element of the sorted taps
template <int K>
__global__ void __launch_bounds__(128) median_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= W) return;
    constexpr int R = K / 2;
    int v[K * K];
    int n = 0;
#pragma unroll
    for (int r = -R; r <= R; ++r)
#pragma unroll
        for (int c = -R; c <= R; ++c) {
            const int y = min(max(i + r, 0), H - 1), x = min(max(j + c, 0), W - 1);
            const int t = in[(size_t)y * W + x];
            int k = n++;
            while (k > 0 && v[k - 1] > t) {
                v[k] = v[k - 1];
                --k;
            }
            v[k] = t;
        }
    out[(size_t)i * W + j] = (uint8_t)v[K * K / 2];
}

static smt_status_t k_median_blur_u8(smt_ctx* ctx, const uint8_t* in, uint8_t* out, int H, int W, int k) {
    dim3 grid(cdiv(W, 128), H);
    switch (k) {
        case 3: SMT_LAUNCH(ctx, median_u8_kernel<3>, grid, 128, 0, in, out, W, H); break;
        case 5: SMT_LAUNCH(ctx, median_u8_kernel<5>, grid, 128, 0, in, out, W, H); break;
        default: return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "median_blur_u8: ksize %d (3 and 5 are supported)", k);
    }
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- FillImageNew
// Per zero pixel (ASW.h:456-501, on the unmodified image): probe columns c, c-1, c-3, c-6, ... for the first non-zero
// value; if that leaves the row, probe c, c+1, c+3, ... where column == W is still read (:489 tests `> col`) -- the first
// pixel of the next row, `past_end` below the last row.  Beyond W the value 0 is recorded; ending exactly on column W with
// a zero there records NOTHING, and because values and pixels are paired by index (:503-509) every later zero pixel then
// receives the value recorded for a later pixel.  flag: 0 = not a zero pixel, 1 = zero pixel with a recorded value,
// 2 = zero pixel without one.
__global__ void __launch_bounds__(256) fin_search_kernel(const uint8_t* __restrict__ img, int H, int W, int past_end, uint8_t* __restrict__ val,
                                                         uint8_t* __restrict__ flag) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W) return;
    if (img[p]) {
        flag[p] = 0;
        return;
    }
    const int i = (int)(p / W), c = (int)(p - (size_t)i * W);
    const uint8_t* row = img + (size_t)i * W;
    int found = -1;
    for (int col = c - 1, off = 1; col >= 0; ++off, col -= off) {  // col == c is the zero pixel itself
        const int v = row[col];
        if (v) {
            found = v;
            break;
        }
    }
    uint8_t f = 1;
    if (found < 0) {
        f = 2;
        for (int col = c + 1, off = 1;; ++off, col += off) {
            if (col > W) {
                found = 0;
                f = 1;
                break;
            }
            const int v = (col == W && i == H - 1) ? past_end : row[col];
            if (v) {
                found = v;
                f = 1;
                break;
            }
            if (col == W) break;  // :485: the loop condition fails before another probe
        }
    }
    val[p] = (uint8_t)max(found, 0);
    flag[p] = f;
}
// per 256-pixel block: number of zero pixels and of recorded values
__global__ void __launch_bounds__(256) fin_count_kernel(const uint8_t* __restrict__ flag, size_t n, int* __restrict__ blockcnt) {
    __shared__ int s_z, s_v;
    if (threadIdx.x == 0) s_z = s_v = 0;
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = p < n ? flag[p] : 0;
    const unsigned bz = __ballot_sync(0xffffffffu, f != 0), bv = __ballot_sync(0xffffffffu, f == 1);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_z, __popc(bz));
        atomicAdd(&s_v, __popc(bv));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        blockcnt[2 * blockIdx.x] = s_z;
        blockcnt[2 * blockIdx.x + 1] = s_v;
    }
}
// single block: exclusive scan of the two per-block counts in place, totals to counts[0..1]
__global__ void __launch_bounds__(1024) fin_scan_kernel(int* __restrict__ blockcnt, int nb, int* __restrict__ counts) {
    __shared__ int s_part[2][1024];
    const int t = threadIdx.x;
    const int per = (nb + 1023) / 1024;
    const int b0 = min(nb, t * per), b1 = min(nb, b0 + per);
    int sz = 0, sv = 0;
    for (int b = b0; b < b1; ++b) {
        sz += blockcnt[2 * b];
        sv += blockcnt[2 * b + 1];
    }
    s_part[0][t] = sz;
    s_part[1][t] = sv;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int a = 0, b = 0;
        if (t >= off) {
            a = s_part[0][t - off];
            b = s_part[1][t - off];
        }
        __syncthreads();
        s_part[0][t] += a;
        s_part[1][t] += b;
        __syncthreads();
    }
    int ez = s_part[0][t] - sz, ev = s_part[1][t] - sv;
    for (int b = b0; b < b1; ++b) {
        const int cz = blockcnt[2 * b], cv = blockcnt[2 * b + 1];
        blockcnt[2 * b] = ez;
        blockcnt[2 * b + 1] = ev;
        ez += cz;
        ev += cv;
    }
    if (t == 1023) {
        counts[0] = s_part[0][t];
        counts[1] = s_part[1][t];
    }
}
// ranks inside the block by ballots; pass 0 compacts the recorded values, pass 1 hands value i to zero pixel i
template <int PASS>
__global__ void __launch_bounds__(256) fin_place_kernel(const uint8_t* __restrict__ flag, const uint8_t* __restrict__ val, size_t n,
                                                        const int* __restrict__ blockcnt, const int* __restrict__ counts,
                                                        uint8_t* __restrict__ compact, uint8_t* __restrict__ img) {
    __shared__ int s_warp[8];
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = p < n ? flag[p] : 0;
    const bool mine = PASS == 0 ? f == 1 : f != 0;
    const unsigned b = __ballot_sync(0xffffffffu, mine);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_warp[w] = __popc(b);
    __syncthreads();
    int before = blockcnt[2 * blockIdx.x + (PASS == 0 ? 1 : 0)];
    for (int k = 0; k < w; ++k) before += s_warp[k];
    const int rank = before + __popc(b & ((1u << lane) - 1u));
    if (!mine) return;
    if (PASS == 0) compact[rank] = val[p];
    else if (rank < counts[1]) img[p] = compact[rank];  // the reference reads past its value vector for the others: left untouched
}

static smt_status_t k_fill_image_new(smt_ctx* ctx, uint8_t* img, int H, int W, int past_end, int* counts_dev) {
    const size_t n = (size_t)H * W;
    const int nb = (int)cdiv((long long)n, 256);
    SMT_ALLOC(ctx, val, n);
    SMT_ALLOC(ctx, flag, n);
    SMT_ALLOC(ctx, compact, n);
    SMT_ALLOC(ctx, bc, (size_t)nb * 2 * sizeof(int));
    SMT_LAUNCH(ctx, fin_search_kernel, nb, 256, 0, img, H, W, past_end, val.as<uint8_t>(), flag.as<uint8_t>());
    SMT_LAUNCH(ctx, fin_count_kernel, nb, 256, 0, flag.as<uint8_t>(), n, bc.as<int>());
    SMT_LAUNCH(ctx, fin_scan_kernel, 1, 1024, 0, bc.as<int>(), nb, counts_dev);
    SMT_LAUNCH(ctx, fin_place_kernel<0>, nb, 256, 0, flag.as<uint8_t>(), val.as<uint8_t>(), n, bc.as<int>(), counts_dev, compact.as<uint8_t>(), img);
    SMT_LAUNCH(ctx, fin_place_kernel<1>, nb, 256, 0, flag.as<uint8_t>(), val.as<uint8_t>(), n, bc.as<int>(), counts_dev, compact.as<uint8_t>(), img);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

smt_status_t smt_normalize_minmax_u8_from_u8(smt_ctx* ctx, const uint8_t* src, int32_t rows, int32_t cols, uint8_t* out, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, src && out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_normalize_minmax_u8_from_u8: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* d = ms.in(src, n);
    uint8_t* o = ms.out(out, n);
    SMT_TRY(ms.check());
    SMT_TRY(k_normalize_u8u8(ctx, d, n, o));
    return ms.finish();
}

smt_status_t smt_filter_speckles_u8(smt_ctx* ctx, uint8_t* img, int32_t rows, int32_t cols, int32_t new_val, int32_t max_speckle_size,
                                    int32_t max_diff, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, img && rows > 0 && cols > 0 && (int64_t)rows * cols < 0x7fffffff && max_diff >= 0, SMT_ERR_INVALID_ARG,
                "smt_filter_speckles_u8: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    uint8_t* d = ms.out(img, n, true);
    SMT_TRY(ms.check());
    {
        StageTimer t(ctx, ST_POST);
        // saturate_cast<uchar>(newVal) like OpenCV's template instantiation; a value no pixel can hold never matches
        SMT_TRY(k_filter_speckles_u8(ctx, d, rows, cols, new_val < 0 ? 0 : (new_val > 255 ? 255 : new_val), max_speckle_size, max_diff));
    }
    return ms.finish();
}

smt_status_t smt_median_blur_u8(smt_ctx* ctx, const uint8_t* in, uint8_t* out, int32_t rows, int32_t cols, int32_t ksize, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, in && out && in != out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_median_blur_u8: bad argument (in place is not supported)");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* d = ms.in(in, n);
    uint8_t* o = ms.out(out, n);
    SMT_TRY(ms.check());
    {
        StageTimer t(ctx, ST_POST);
        SMT_TRY(k_median_blur_u8(ctx, d, o, rows, cols, ksize));
    }
    return ms.finish();
}

smt_status_t smt_fill_image_new(smt_ctx* ctx, uint8_t* disp, int32_t rows, int32_t cols, int32_t past_end, int32_t* n_unassigned,
                                smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && rows > 0 && cols > 0 && (int64_t)rows * cols < 0x7fffffff, SMT_ERR_INVALID_ARG, "smt_fill_image_new: bad argument");
    // a successful search to the left parks the column at 0xffff (:480) so that the search to the right is skipped
    SMT_REQUIRE(ctx, cols <= 0xffff, SMT_ERR_UNSUPPORTED, "smt_fill_image_new: more than 65535 columns (the reference's 0xffff sentinel)");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    uint8_t* d = ms.out(disp, n, true);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, counts, 2 * sizeof(int));
    {
        StageTimer t(ctx, ST_POST);
        SMT_TRY(k_fill_image_new(ctx, d, rows, cols, past_end & 0xff, counts.as<int>()));
    }
    if (n_unassigned) {
        SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, counts.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *n_unassigned = ctx->h_pinned[0] - ctx->h_pinned[1];
    }
    return ms.finish();
}

smt_status_t smt_asw_display_chain(smt_ctx* ctx, const float* disp_left, const float* disp_right, int32_t rows, int32_t cols,
                                   uint8_t* speckle_img, uint8_t* filled_img, uint8_t* median_img, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp_left && disp_right && median_img && rows > 0 && cols > 0 && (int64_t)rows * cols < 0x7fffffff, SMT_ERR_INVALID_ARG,
                "smt_asw_display_chain: bad argument");
    SMT_REQUIRE(ctx, cols <= 0xffff, SMT_ERR_UNSUPPORTED, "smt_asw_display_chain: more than 65535 columns");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* dl = ms.in(disp_left, n);
    const float* dr = ms.in(disp_right, n);
    uint8_t* o_spk = ms.out(speckle_img, n);
    uint8_t* o_fill = ms.out(filled_img, n);
    uint8_t* o_med = ms.out(median_img, n);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, a, n);
    SMT_ALLOC(ctx, b, n);
    SMT_ALLOC(ctx, counts, 2 * sizeof(int));
    SMT_TRY(k_asw_crosscheck(ctx, dl, dr, rows, cols, a.as<uint8_t>()));  // ASWeight.cpp:66
    StageTimer t(ctx, ST_POST);
    SMT_TRY(k_normalize_u8u8(ctx, a.as<uint8_t>(), n, b.as<uint8_t>()));                    // :69 (:72 converts CV_8U to CV_8U)
    SMT_TRY(k_filter_speckles_u8(ctx, b.as<uint8_t>(), rows, cols, 0, 40, 2));             // :73
    SMT_TRY(k_median_blur_u8(ctx, b.as<uint8_t>(), a.as<uint8_t>(), rows, cols, 5));       // :74
    if (o_spk) SMT_CUDA(ctx, cudaMemcpyAsync(o_spk, a.p, n, cudaMemcpyDeviceToDevice, ctx->stream));  // "speckle.png"
    SMT_TRY(k_fill_image_new(ctx, a.as<uint8_t>(), rows, cols, 0, counts.as<int>()));      // :76
    if (o_fill) SMT_CUDA(ctx, cudaMemcpyAsync(o_fill, a.p, n, cudaMemcpyDeviceToDevice, ctx->stream));  // "filled.png"
    SMT_TRY(k_median_blur_u8(ctx, a.as<uint8_t>(), o_med, rows, cols, 3));                 // :78
    return ms.finish();
}

}  // extern "C"
