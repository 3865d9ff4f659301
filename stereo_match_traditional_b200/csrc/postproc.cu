// Disparity post-processing: left-right consistency (both reference variants), speckle removal, median filter.
//
// Replaces LeftRightConsistency, LeftAndRightConsistency, RemoveSpeckles, MedianFilter
// (AD-CensusV1/PostProcessing.h:72-135, :10-70, :250-311, :314-344).
#include <cmath>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------- LR check
// The reference walks pixels in row-major order and invalidates IN PLACE, so the occlusion/mismatch test of a later
// pixel can read a +inf written earlier in the same row (:112).  Parallel form: phase 1 decides validity from the
// original values only; phase 2 classifies, substituting +inf for an already-invalidated pixel to the LEFT of the
// current one; phase 3 writes +inf and compacts the two (row, col) lists in row-major order.
enum { LR_VALID = -1, LR_MISMATCH = -2 };  // code >= 0: column to classify against

__global__ void __launch_bounds__(256) lr_phase1_kernel(const float* __restrict__ dl, const float* __restrict__ dr, int W, int H,
                                                        float gate, int variant_and, int* __restrict__ code) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
    const float d = dl[p];
    int c = LR_MISMATCH;
    if (!variant_and && d == INFINITY) {  // :90-93
        code[p] = LR_MISMATCH;
        return;
    }
    const int col_right = (int)((double)((float)j - d) + 0.5);  // :96 / :25 (int - float is float; + 0.5 is double)
    if (col_right >= 0 && col_right < W) {
        const float r = dr[(size_t)i * W + col_right];
        const float diff = fabsf(d - r);
        const bool bad = variant_and ? (diff >= gate) : (diff > gate);  // :33 vs :103
        if (bad) {
            const int col_rl = (int)((double)((float)col_right + r) + 0.5);  // :110 / :40
            c = (col_rl > 0 && col_rl < W) ? col_rl : LR_MISMATCH;
        } else {
            c = LR_VALID;
        }
    }
    code[p] = c;
}

// cls: 0 valid, 1 occlusion, 2 mismatch
__global__ void __launch_bounds__(256) lr_phase2_kernel(const float* __restrict__ dl, const int* __restrict__ code, int W, int H,
                                                        int variant_and, unsigned char* __restrict__ cls) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
    const int c = code[p];
    unsigned char k = 0;
    if (c == LR_MISMATCH) {
        k = 2;
    } else if (c >= 0) {
        const size_t q = (size_t)i * W + c;
        float v = dl[q];
        if (!variant_and && c < j && code[q] != LR_VALID) v = INFINITY;  // already overwritten in the reference (:112,:125)
        k = (v > dl[p]) ? 1 : 2;
    }
    cls[p] = k;
}

__global__ void __launch_bounds__(256) lr_count_kernel(const unsigned char* __restrict__ cls, size_t n, int* __restrict__ blockcnt) {
    __shared__ int s_occ, s_mis;
    if (threadIdx.x == 0) s_occ = s_mis = 0;
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = p < n ? cls[p] : 0;
    const unsigned bo = __ballot_sync(0xffffffffu, k == 1), bm = __ballot_sync(0xffffffffu, k == 2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_occ, __popc(bo));
        atomicAdd(&s_mis, __popc(bm));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        blockcnt[2 * blockIdx.x] = s_occ;
        blockcnt[2 * blockIdx.x + 1] = s_mis;
    }
}

// single block: exclusive scan of the per-block counts (in place), totals to counts[0..1]
__global__ void __launch_bounds__(1024) lr_scan_kernel(int* __restrict__ blockcnt, int nb, int* __restrict__ counts) {
    __shared__ int s_part[2][1024];
    const int t = threadIdx.x;
    const int per = (nb + 1023) / 1024;
    const int b0 = t * per, b1 = min(nb, b0 + per);
    int so = 0, sm = 0;
    for (int b = b0; b < b1; ++b) {
        so += blockcnt[2 * b];
        sm += blockcnt[2 * b + 1];
    }
    s_part[0][t] = so;
    s_part[1][t] = sm;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
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
    int eo = s_part[0][t] - so, em = s_part[1][t] - sm;
    for (int b = b0; b < b1; ++b) {
        const int co = blockcnt[2 * b], cm = blockcnt[2 * b + 1];
        blockcnt[2 * b] = eo;
        blockcnt[2 * b + 1] = em;
        eo += co;
        em += cm;
    }
    if (t == 1023) {
        counts[0] = s_part[0][1023];
        counts[1] = s_part[1][1023];
    }
}

__global__ void __launch_bounds__(256) lr_apply_kernel(const unsigned char* __restrict__ cls, const int* __restrict__ blockoff, int W,
                                                       int H, int variant_and, float* dl_inout,
                                                       const float* dl_orig, float* __restrict__ last,
                                                       int* __restrict__ occ, int* __restrict__ mis) {
    __shared__ int s_wo[8], s_wm[8];
    const size_t n = (size_t)W * H;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = p < n ? cls[p] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bo = __ballot_sync(0xffffffffu, k == 1), bm = __ballot_sync(0xffffffffu, k == 2);
    if (lane == 0) {
        s_wo[warp] = __popc(bo);
        s_wm[warp] = __popc(bm);
    }
    __syncthreads();
    int base_o = blockoff[2 * blockIdx.x], base_m = blockoff[2 * blockIdx.x + 1];
    for (int w = 0; w < warp; ++w) {
        base_o += s_wo[w];
        base_m += s_wm[w];
    }
    if (p >= n) return;
    const unsigned lower = (1u << lane) - 1u;
    const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
    if (k == 1 && occ) {
        const int r = base_o + __popc(bo & lower);
        occ[2 * r] = i;
        occ[2 * r + 1] = j;
    }
    if (k == 2 && mis) {
        const int r = base_m + __popc(bm & lower);
        mis[2 * r] = i;
        mis[2 * r + 1] = j;
    }
    if (variant_and) {
        last[p] = k ? 0.f : dl_orig[p];  // :56,:60,:65
    } else if (k) {
        dl_inout[p] = INFINITY;  // :125,:130 (already +inf inputs stay +inf)
    }
}

smt_status_t k_lrcheck(smt_ctx* ctx, float* dispL, const float* dispR, int W, int H, float gate, bool variant_and, float* last,
                       int32_t* occ, int32_t* mis, int32_t* counts_dev) {
    StageTimer t(ctx, ST_POST);
    const size_t n = (size_t)W * H;
    const int nb = (int)cdiv((long long)n, 256);
    SMT_ALLOC(ctx, code, n * sizeof(int));
    SMT_ALLOC(ctx, cls, n);
    SMT_ALLOC(ctx, bc, (size_t)nb * 2 * sizeof(int));
    const int va = variant_and ? 1 : 0;
    SMT_LAUNCH(ctx, lr_phase1_kernel, nb, 256, 0, dispL, dispR, W, H, gate, va, code.as<int>());
    SMT_LAUNCH(ctx, lr_phase2_kernel, nb, 256, 0, dispL, code.as<int>(), W, H, va, cls.as<unsigned char>());
    SMT_LAUNCH(ctx, lr_count_kernel, nb, 256, 0, cls.as<unsigned char>(), n, bc.as<int>());
    SMT_LAUNCH(ctx, lr_scan_kernel, 1, 1024, 0, bc.as<int>(), nb, counts_dev);
    SMT_LAUNCH(ctx, lr_apply_kernel, nb, 256, 0, cls.as<unsigned char>(), bc.as<int>(), W, H, va, dispL, dispL, last, occ, mis);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- speckles
// The reference's BFS (:260-310) finds the connected components of the undirected graph {8-neighbours, both valid,
// |delta| <= diff}; membership does not depend on visiting order, so a lock-free union-find gives the same
// components.  Components smaller than min_area are overwritten with (float)invalid_val.
__device__ __forceinline__ int uf_find(volatile int* L, int a) {
    int b = L[a];
    while (b != a) {
        a = b;
        b = L[a];
    }
    return a;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// edge predicate of the reference's region growing (:290-292): both valid, |delta| <= diff
__device__ __forceinline__ bool spk_conn(float a, float b, float invalid, float diff) {
    return a != invalid && b != invalid && fabsf(a - b) <= diff;
}

// Stage 1: one block per row.  label = start index of the pixel's horizontal run (pointer jumping in shared
// memory, log2(W) rounds), -1 for invalid pixels.  Rows are the long direction, so this removes the long chains a
// naive per-pixel union-find would have to walk.
__global__ void __launch_bounds__(256) spk_rows_kernel(const float* __restrict__ disp, int W, float invalid, float diff,
                                                       int* __restrict__ label, int* __restrict__ size) {
    extern __shared__ int s_l[];  // W ints
    const int i = blockIdx.x;
    const float* row = disp + (size_t)i * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        const float v = row[j];
        int l = j;
        if (v == invalid) l = -1;  // :262 (float == (float)int)
        else if (j > 0 && spk_conn(v, row[j - 1], invalid, diff)) l = j - 1;
        s_l[j] = l;
        size[(size_t)i * W + j] = 0;
    }
    __syncthreads();
    for (int span = 1; span < W; span <<= 1) {
        // monotone pointer jumping: reading a value another thread is updating only ever yields an ancestor
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

// Stage 2: link runs of adjacent rows.  A (pixel, neighbour-below) edge is skipped when edges further to the left
// already imply it (first-contact rule), so large regions issue a handful of unions instead of three per pixel.
__global__ void __launch_bounds__(256) spk_merge_kernel(const float* __restrict__ disp, int W, int H, float invalid, float diff,
                                                        int* __restrict__ label) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
    if (i + 1 >= H) return;
    const float v = disp[p];
    if (v == invalid) return;
    const float* r0 = disp + (size_t)i * W;
    const float* r1 = r0 + W;
    const float vl = j > 0 ? r0[j - 1] : invalid, vr = j + 1 < W ? r0[j + 1] : invalid;
    const float sw = j > 0 ? r1[j - 1] : invalid, ss = r1[j], se = j + 1 < W ? r1[j + 1] : invalid;
    const bool c_l = spk_conn(v, vl, invalid, diff), c_r = spk_conn(v, vr, invalid, diff);
    const bool c_s = spk_conn(v, ss, invalid, diff), c_sw = spk_conn(v, sw, invalid, diff), c_se = spk_conn(v, se, invalid, diff);
    const size_t q = p + W;
    if (c_s && !(c_l && spk_conn(ss, sw, invalid, diff) && spk_conn(vl, sw, invalid, diff))) uf_union(label, (int)p, (int)q);
    if (c_sw && !(c_l && spk_conn(vl, sw, invalid, diff)) && !(c_s && spk_conn(ss, sw, invalid, diff)))
        uf_union(label, (int)p, (int)(q - 1));
    if (c_se && !(c_s && spk_conn(ss, se, invalid, diff)) && !(c_r && spk_conn(vr, se, invalid, diff)))
        uf_union(label, (int)p, (int)(q + 1));
}
__global__ void __launch_bounds__(256) spk_count_kernel(int* __restrict__ label, size_t n, int* __restrict__ size) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int root = -1;
    if (p < n && label[p] >= 0) {
        root = uf_find(label, (int)p);
        label[p] = root;  // roots keep label[root]==root, so concurrent finds stay valid
    }
    // one atomic per distinct root per warp (large regions would otherwise serialise on a single address)
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (root >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&size[root], __popc(peers));
}
__global__ void __launch_bounds__(256) spk_apply_kernel(const int* __restrict__ label, const int* __restrict__ size, size_t n,
                                                        unsigned min_area, float invalid, float* __restrict__ disp) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l = label[p];
    if (l < 0) return;
    if ((unsigned)size[l] < min_area) disp[p] = invalid;  // :304-307
}

smt_status_t k_remove_speckles(smt_ctx* ctx, float* disp, int W, int H, int diff, uint32_t min_area, int invalid_val) {
    StageTimer t(ctx, ST_POST);
    const size_t n = (size_t)W * H;
    const unsigned nb = cdiv((long long)n, 256);
    SMT_ALLOC(ctx, label, n * sizeof(int));
    SMT_ALLOC(ctx, size, n * sizeof(int));
    const float inv = (float)invalid_val;
    if ((size_t)W * sizeof(int) > 200 * 1024) return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "remove_speckles: width %d too large", W);
    if ((size_t)W * sizeof(int) > 48 * 1024)
        SMT_CUDA(ctx, cudaFuncSetAttribute(spk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(W * sizeof(int))));
    SMT_LAUNCH(ctx, spk_rows_kernel, H, 256, (size_t)W * sizeof(int), disp, W, inv, (float)diff, label.as<int>(), size.as<int>());
    SMT_LAUNCH(ctx, spk_merge_kernel, nb, 256, 0, disp, W, H, inv, (float)diff, label.as<int>());
    SMT_LAUNCH(ctx, spk_count_kernel, nb, 256, 0, label.as<int>(), n, size.as<int>());
    SMT_LAUNCH(ctx, spk_apply_kernel, nb, 256, 0, label.as<int>(), size.as<int>(), n, min_area, inv, disp);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- median
template <int K>
__global__ void __launch_bounds__(128) median_kernel(const float* __restrict__ in, float* __restrict__ out, int W, int H) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= W) return;
    constexpr int R = K / 2;
    float v[K * K];
    int n = 0;
#pragma unroll
    for (int r = -R; r <= R; ++r)
#pragma unroll
        for (int c = -R; c <= R; ++c) {
            const int y = i + r, x = j + c;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                // insertion into the sorted prefix
                const float t = in[(size_t)y * W + x];
                int k = n++;
                while (k > 0 && v[k - 1] > t) {
                    v[k] = v[k - 1];
                    --k;
                }
                v[k] = t;
            }
        }
    out[(size_t)i * W + j] = v[n / 2];  // :341
}

smt_status_t k_median(smt_ctx* ctx, const float* in, float* out, int W, int H, int k) {
    StageTimer t(ctx, ST_POST);
    dim3 grid(cdiv(W, 128), H);
    switch (k) {
        case 1: SMT_LAUNCH(ctx, median_kernel<1>, grid, 128, 0, in, out, W, H); break;
        case 3: SMT_LAUNCH(ctx, median_kernel<3>, grid, 128, 0, in, out, W, H); break;
        case 5: SMT_LAUNCH(ctx, median_kernel<5>, grid, 128, 0, in, out, W, H); break;
        case 7: SMT_LAUNCH(ctx, median_kernel<7>, grid, 128, 0, in, out, W, H); break;
        case 9: SMT_LAUNCH(ctx, median_kernel<9>, grid, 128, 0, in, out, W, H); break;
        default: return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "median: window size %d (odd sizes 1..9 are supported)", k);
    }
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI
static smt_status_t lrcheck_common(smt_ctx* ctx, float* dl_inout, const float* dl_const, const float* dr, float* last, int cols,
                                   int rows, float gate, bool variant_and, int32_t* occ, int32_t* n_occ, int32_t* mis,
                                   int32_t* n_mis, smt_memspace_t space) {
    SMT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    float* dl = nullptr;
    if (variant_and)
        dl = const_cast<float*>(ms.in(dl_const, n));
    else
        dl = ms.out(dl_inout, n, /*also_in=*/true);
    const float* ddr = ms.in(dr, n);
    float* dlast = ms.out(last, n);
    int32_t* docc = ms.out(occ, 2 * n);
    int32_t* dmis = ms.out(mis, 2 * n);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, counts, 2 * sizeof(int));
    SMT_TRY(k_lrcheck(ctx, dl, ddr, cols, rows, gate, variant_and, dlast, docc, dmis, counts.as<int>()));
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, counts.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_occ) *n_occ = ctx->h_pinned[0];
    if (n_mis) *n_mis = ctx->h_pinned[1];
    return ms.finish();
}

extern "C" {

smt_status_t smt_lrcheck(smt_ctx* ctx, float* disp_left, const float* disp_right, int32_t cols, int32_t rows, int32_t gate,
                         int32_t* occ, int32_t* n_occ, int32_t* mis, int32_t* n_mis, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp_left && disp_right && cols > 0 && rows > 0, SMT_ERR_INVALID_ARG, "smt_lrcheck: bad argument");
    return lrcheck_common(ctx, disp_left, nullptr, disp_right, nullptr, cols, rows, (float)gate, false, occ, n_occ, mis, n_mis,
                          space);
}

smt_status_t smt_lrcheck_and(smt_ctx* ctx, const float* disp_left, const float* disp_right, float* last, int32_t cols,
                             int32_t rows, float gate, int32_t* occ, int32_t* n_occ, int32_t* mis, int32_t* n_mis,
                             smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp_left && disp_right && last && cols > 0 && rows > 0, SMT_ERR_INVALID_ARG, "smt_lrcheck_and: bad argument");
    return lrcheck_common(ctx, nullptr, disp_left, disp_right, last, cols, rows, gate, true, occ, n_occ, mis, n_mis, space);
}

smt_status_t smt_remove_speckles(smt_ctx* ctx, float* disp, int32_t cols, int32_t rows, int32_t diff, uint32_t min_area,
                                 int32_t invalid_val, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && cols > 0 && rows > 0, SMT_ERR_INVALID_ARG, "smt_remove_speckles: bad argument");
    SMT_CUDA(ctx, cudaSetDevice(ctx->device));
    Marshal ms(ctx, space);
    float* d = ms.out(disp, (size_t)rows * cols, true);
    SMT_TRY(ms.check());
    SMT_TRY(k_remove_speckles(ctx, d, cols, rows, diff, min_area, invalid_val));
    return ms.finish();
}

smt_status_t smt_median(smt_ctx* ctx, const float* in, float* out, int32_t cols, int32_t rows, int32_t wnd_size,
                        smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, in && out && cols > 0 && rows > 0, SMT_ERR_INVALID_ARG, "smt_median: bad argument");
    SMT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* di = ms.in(in, n);
    float* dout = ms.out(out, n);
    SMT_TRY(ms.check());
    SMT_TRY(k_median(ctx, di, dout, cols, rows, wnd_size));
    return ms.finish();
}

}  // extern "C"
