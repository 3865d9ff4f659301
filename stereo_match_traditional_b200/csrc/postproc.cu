// Disparity post-processing: left-right consistency (both reference variants), speckle removal, median filter.
//
// Replaces LeftRightConsistency, LeftAndRightConsistency, RemoveSpeckles, MedianFilter
// (AD-CensusV1/PostProcessing.h:72-135, :10-70, :250-311, :314-344).
#include <cmath>

#include "common.cuh"
#include "unionfind.cuh"

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
    } else if (k && dl_inout) {
        dl_inout[p] = INFINITY;  // :125,:130 (already +inf inputs stay +inf)
    }
}

smt_status_t k_lrcheck(smt_ctx* ctx, float* dispL, const float* dispR, int W, int H, float gate, bool variant_and, float* last,
                       int32_t* occ, int32_t* mis, int32_t* counts_dev, uint8_t* cls_out) {
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
    if (cls_out) SMT_CUDA(ctx, cudaMemcpyAsync(cls_out, cls.p, n, cudaMemcpyDeviceToDevice, ctx->stream));
    return SMT_OK;
}

// class map (1 = occlusion, 2 = mismatch) -> the two (row, col) lists in raster order; counts_dev[0..1] = their lengths
smt_status_t k_class_map_to_lists(smt_ctx* ctx, const unsigned char* cls, int W, int H, int32_t* occ, int32_t* mis, int32_t* counts_dev) {
    const size_t n = (size_t)W * H;
    const int nb = (int)cdiv((long long)n, 256);
    SMT_ALLOC(ctx, bc, (size_t)nb * 2 * sizeof(int));
    SMT_LAUNCH(ctx, lr_count_kernel, nb, 256, 0, cls, n, bc.as<int>());
    SMT_LAUNCH(ctx, lr_scan_kernel, 1, 1024, 0, bc.as<int>(), nb, counts_dev);
    SMT_LAUNCH(ctx, lr_apply_kernel, nb, 256, 0, cls, bc.as<int>(), W, H, 0, (float*)nullptr, (const float*)nullptr, (float*)nullptr, occ, mis);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- speckles
// The reference's BFS (:260-310) finds the connected components of the undirected graph {8-neighbours, both valid,
// |delta| <= diff}; membership does not depend on visiting order, so a lock-free union-find gives the same
// components.  Components smaller than min_area are overwritten with (float)invalid_val.
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
    SMT_LAUNCH(ctx, uf_count_kernel, nb, 256, 0, label.as<int>(), n, size.as<int>());
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

// ---------------------------------------------------------------------------------------------- hole filling
// FillTheHole (AD-CensusV1/PostProcessing.h:156-248) exactly as written; the function is never called by the reference's
// main and is kept for API completeness.  What "as written" means:
//  * width = row and height = col (:158-159): pixel (y, x) of a list addresses disp[y*row + x], y < col, x < row;
//  * "invalid" is the value 0xffff (:184, :215), not the +inf LeftRightConsistency writes, so +inf is collected as valid;
//  * the direction table switches from angle1 to angle2 at the first target (in list order) whose y equals height/2 and
//    never switches back, also across the three phases (:198-200) -> `first` / `switched` below;
//  * a target without any hit is filled with 0 (value-initialised fill_disps, :177, :239-244);
//  * a phase reads the map of the previous phase only (fills are written after the whole list was visited, :239), so a
//    phase is one parallel kernel; duplicates in a list resolve to the LAST entry like the sequential write loop;
//  * phase 3 replaces the mismatch list by all 0xffff pixels in raster order (:179-189) and is skipped when the mismatch
//    list was empty (:175).  The reference sizes fill_disps before that replacement, so it overflows the heap when
//    there are more 0xffff pixels than mismatches; here that case simply fills all of them.
struct FillDirs {
    float sina[2][8], cosa[2][8];
};

// every (y, x) of a caller-supplied list must address the frame: checked before anything reads or writes through it
__global__ void __launch_bounds__(256) fill_validate_kernel(const int* __restrict__ trg, int ntrg, int width, int height,
                                                            int* __restrict__ bad) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ntrg) return;
    const int y = trg[2 * n], x = trg[2 * n + 1];
    if (y < 0 || y >= height || x < 0 || x >= width) *bad = 1;
}

__global__ void __launch_bounds__(256) fill_scan_kernel(const int* __restrict__ trg, int ntrg, int width, int height,
                                                        int* __restrict__ first, int* __restrict__ owner) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ntrg) return;
    const int y = trg[2 * n], x = trg[2 * n + 1];
    if (y == height / 2) atomicMin(first, n);
    atomicMax(&owner[(size_t)y * width + x], n);
}

__global__ void __launch_bounds__(256) fill_collect_kernel(const float* __restrict__ disp, const int* __restrict__ trg, int ntrg,
                                                           int width, int height, int max_len, int phase, float invalid, FillDirs dirs,
                                                           const int* __restrict__ first, const int* __restrict__ switched,
                                                           float* __restrict__ fill) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ntrg) return;
    const int y = trg[2 * n], x = trg[2 * n + 1];
    const int tab = (*switched || n >= *first) ? 1 : 0;
    float coll[8];
    int nc = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const float sina = dirs.sina[tab][s], cosa = dirs.cosa[tab][s];
        for (int m = 1; m < max_len; ++m) {
            const int yy = (int)lroundf(__fadd_rn((float)y, __fmul_rn((float)m, sina)));  // :209
            const int xx = (int)lroundf(__fadd_rn((float)x, __fmul_rn((float)m, cosa)));
            if (yy < 0 || yy >= height || xx < 0 || xx >= width) break;
            const float d = disp[(size_t)yy * width + xx];
            if (d != invalid) {
                // insertion into the sorted prefix (std::sort of <= 8 values, :226)
                int k = nc++;
#pragma unroll
                for (int q = 7; q > 0; --q)
                    if (q <= k && coll[q - 1] > d) {
                        coll[q] = coll[q - 1];
                        k = q - 1;
                    }
                coll[k] = d;
                break;
            }
        }
    }
    float v = 0.f;
    if (nc) v = phase == 0 ? coll[nc > 1 ? 1 : 0] : coll[nc / 2];  // :230-241
    fill[n] = v;
}

__global__ void __launch_bounds__(256) fill_write_kernel(float* __restrict__ disp, const int* __restrict__ trg, int ntrg, int width,
                                                         const float* __restrict__ fill, int* __restrict__ owner,
                                                         int* __restrict__ first, int* __restrict__ switched) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0 && *first != 0x7fffffff) *switched = 1;
    if (n >= ntrg) return;
    const size_t p = (size_t)trg[2 * n] * width + trg[2 * n + 1];
    if (owner[p] == n) {
        disp[p] = fill[n];
        owner[p] = -1;
    }
}

__global__ void __launch_bounds__(256) fill_mark_kernel(const float* __restrict__ disp, size_t n, float invalid,
                                                        unsigned char* __restrict__ cls) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) cls[p] = disp[p] == invalid ? 2 : 0;
}

// variant SMT_FILL_FIXED: what the function evidently means -- width = col, height = row (list entries are the (row, col)
// pairs LeftRightConsistency emits), invalid = +inf (what LeftRightConsistency writes); everything else unchanged.
smt_status_t k_fill_the_hole(smt_ctx* ctx, float* disp, int row, int col, int disp_range, const int32_t* occ, int n_occ, int32_t* mis,
                             int* n_mis_inout, int variant) {
    StageTimer t(ctx, ST_POST);
    const bool fixed = variant == SMT_FILL_FIXED;
    const int width = fixed ? col : row, height = fixed ? row : col;
    const float invalid = fixed ? INFINITY : 65535.f;
    const size_t n = (size_t)row * col;
    FillDirs dirs;
    {
        const float pi = 3.1415926f;  // :164-166, evaluated in float like the reference's initialisers
        const float a1[8] = {pi, 3 * pi / 4, pi / 2, pi / 4, 0, 7 * pi / 4, 3 * pi / 2, 5 * pi / 4};
        const float a2[8] = {pi, 5 * pi / 4, 3 * pi / 2, 7 * pi / 4, 0, pi / 4, pi / 2, 3 * pi / 4};
        for (int s = 0; s < 8; ++s) {
            dirs.sina[0][s] = float(sin(a1[s]));  // :206-207, host libm like the LUTs of the cost kernel
            dirs.cosa[0][s] = float(cos(a1[s]));
            dirs.sina[1][s] = float(sin(a2[s]));
            dirs.cosa[1][s] = float(cos(a2[s]));
        }
    }
    SMT_ALLOC(ctx, owner, n * sizeof(int));
    SMT_ALLOC(ctx, fill, n * sizeof(float));
    SMT_ALLOC(ctx, st, 4 * sizeof(int));  // first, switched, bad, pad
    int* first = st.as<int>();
    int* switched = first + 1;
    int* bad = first + 2;
    SMT_CUDA(ctx, cudaMemsetAsync(owner.p, 0xff, n * sizeof(int), ctx->stream));
    SMT_CUDA(ctx, cudaMemsetAsync(st.p, 0, 4 * sizeof(int), ctx->stream));
    int n_mis = *n_mis_inout;
    // out-of-frame entries (in the as-written variant the axes are swapped, so an ordinary LeftRightConsistency list of a
    // frame with rows > cols has them) are refused before the map is touched
    if (n_occ > 0) SMT_LAUNCH(ctx, fill_validate_kernel, cdiv(n_occ, 256), 256, 0, occ, n_occ, width, height, bad);
    if (n_mis > 0) SMT_LAUNCH(ctx, fill_validate_kernel, cdiv(n_mis, 256), 256, 0, mis, n_mis, width, height, bad);
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned[0])
        return smt_fail(ctx, SMT_ERR_INVALID_ARG,
                        "smt_fill_the_hole: list entry outside the frame (as written: y < col, x < row, PostProcessing.h:158-159)");
    for (int k = 0; k < 3; ++k) {
        const int32_t* trg = k == 0 ? occ : mis;
        int ntrg = k == 0 ? n_occ : n_mis;
        if (ntrg == 0) continue;  // :175 (phase 3 looks at the OLD mismatch list)
        if (k == 2) {
            const int nb = (int)cdiv((long long)n, 256);
            SMT_ALLOC(ctx, cls, n);
            SMT_ALLOC(ctx, bc, (size_t)nb * 2 * sizeof(int));
            SMT_ALLOC(ctx, counts, 2 * sizeof(int));
            SMT_LAUNCH(ctx, fill_mark_kernel, nb, 256, 0, disp, n, invalid, cls.as<unsigned char>());
            SMT_LAUNCH(ctx, lr_count_kernel, nb, 256, 0, cls.as<unsigned char>(), n, bc.as<int>());
            SMT_LAUNCH(ctx, lr_scan_kernel, 1, 1024, 0, bc.as<int>(), nb, counts.as<int>());
            SMT_LAUNCH(ctx, lr_apply_kernel, nb, 256, 0, cls.as<unsigned char>(), bc.as<int>(), width, height, 0, (float*)nullptr,
                       (const float*)nullptr, (float*)nullptr, (int*)nullptr, mis);
            SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, counts.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ntrg = n_mis = ctx->h_pinned[1];
            if (ntrg == 0) break;
        }
        const int nbt = (int)cdiv(ntrg, 256);
        const int init_first = 0x7fffffff;
        SMT_CUDA(ctx, cudaMemcpyAsync(first, &init_first, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        SMT_LAUNCH(ctx, fill_scan_kernel, nbt, 256, 0, trg, ntrg, width, height, first, owner.as<int>());
        SMT_LAUNCH(ctx, fill_collect_kernel, nbt, 256, 0, disp, trg, ntrg, width, height, disp_range, k, invalid, dirs, first, switched,
                   fill.as<float>());
        SMT_LAUNCH(ctx, fill_write_kernel, nbt, 256, 0, disp, trg, ntrg, width, fill.as<float>(), owner.as<int>(), first, switched);
    }
    *n_mis_inout = n_mis;
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI
static smt_status_t lrcheck_common(smt_ctx* ctx, float* dl_inout, const float* dl_const, const float* dr, float* last, int cols,
                                   int rows, float gate, bool variant_and, int32_t* occ, int32_t* n_occ, int32_t* mis,
                                   int32_t* n_mis, smt_memspace_t space) {
    DeviceGuard dev_guard(ctx->device);
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
    DeviceGuard dev_guard(ctx->device);
    Marshal ms(ctx, space);
    float* d = ms.out(disp, (size_t)rows * cols, true);
    SMT_TRY(ms.check());
    SMT_TRY(k_remove_speckles(ctx, d, cols, rows, diff, min_area, invalid_val));
    return ms.finish();
}

smt_status_t smt_fill_the_hole(smt_ctx* ctx, float* disp, int32_t row, int32_t col, int32_t disp_range, const int32_t* occ,
                               int32_t n_occ, int32_t* mis, int32_t* n_mis, int32_t variant, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && row > 0 && col > 0 && n_occ >= 0 && n_mis && *n_mis >= 0 && (occ || n_occ == 0) && mis &&
                         (variant == SMT_FILL_AS_WRITTEN || variant == SMT_FILL_FIXED),
                SMT_ERR_INVALID_ARG, "smt_fill_the_hole: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)row * col;
    SMT_REQUIRE(ctx, (size_t)n_occ <= n && (size_t)*n_mis <= n, SMT_ERR_INVALID_ARG, "smt_fill_the_hole: list longer than the image");
    Marshal ms(ctx, space);
    float* d = ms.out(disp, n, true);
    const int32_t* docc = n_occ ? ms.in(occ, 2 * (size_t)n_occ) : nullptr;
    int32_t* dmis = ms.out(mis, 2 * n, true);
    SMT_TRY(ms.check());
    int nm = *n_mis;
    SMT_TRY(k_fill_the_hole(ctx, d, row, col, disp_range, docc, n_occ, dmis, &nm, variant));
    *n_mis = nm;
    return ms.finish();
}

smt_status_t smt_median(smt_ctx* ctx, const float* in, float* out, int32_t cols, int32_t rows, int32_t wnd_size,
                        smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, in && out && cols > 0 && rows > 0, SMT_ERR_INVALID_ARG, "smt_median: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* di = ms.in(in, n);
    float* dout = ms.out(out, n);
    SMT_TRY(ms.check());
    SMT_TRY(k_median(ctx, di, dout, cols, rows, wnd_size));
    return ms.finish();
}

}  // extern "C"
