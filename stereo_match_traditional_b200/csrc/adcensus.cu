// AD + 9x7 census matching cost and first-minimum WTA.
//
// Replaces AD_Census::{ComputeAD, ComputeADRight, ComputeCensus9x7, ComputeCensus9x7Right, ComputeADcensus,
// ComputeADcensusRight, WTA} (AD-CensusV1/AD-Census.h:75-380).  The reference recomputes both 63-tap census strings
// for every (pixel, d); here every string is computed ONCE as a packed uint64 per pixel (five small maps), and the
// cost kernel is a pure streaming write of the [H][W][D] volume: one XOR + POPC + two table look-ups per cell.
//
// Census maps (bit order = the reference's MSB-first shift-in order, AD-Census.h:167-191):
//   CL, CR  own census of the left / right image, out-of-image taps 0                         [H][W]
//   M       validity mask of the anchor pixel: tap (r,c) valid iff (i+r, j+c) is inside       [H][W]
//           (the reference tests validity in the ANCHOR's frame for both strings, :173,:238)
//   CRx     right-image string seen from virtual column x' = j-d in [-3, W-1]: centre R(i,max(x',0)),
//           tap R(i+r, max(x'+c,0)) (:159-164,:177-182); constant for x' <= -3                [H][W+3]
//   CLx     left-image string seen from x' = j+d in [0, W+3]: centre L(i,min(x',W-1)), tap L(i+r, x'+c) but
//           column 0 (not W-1) when x'+c >= W (:224-229,:242-247); constant for x' >= W+3     [H][W+4]
// cost(i,j,d) = lutAD[|own - other|] + lutC[popc((Cown(i,j) ^ Cext(i,x')) & M(i,j))]   (:287-289), lut* evaluated on the
// host with the same libm expf the reference calls, so the volume is bit-identical.
#include <cmath>

#include "common.cuh"

size_t census_maps_bytes(int H, int W) {
    return sizeof(uint64_t) * ((size_t)H * W * 3 + (size_t)H * (W + 3) + (size_t)H * (W + 4));
}
void census_maps_carve(void* base, int H, int W, CensusMaps* m) {
    uint64_t* p = static_cast<uint64_t*>(base);
    m->CL = p;
    p += (size_t)H * W;
    m->CR = p;
    p += (size_t)H * W;
    m->M = p;
    p += (size_t)H * W;
    m->CRx = p;
    p += (size_t)H * (W + 3);
    m->CLx = p;
}

enum { CM_OWN = 0, CM_RIGHT_EXT = 1, CM_LEFT_EXT = 2, CM_MASK = 3 };

// the 63-bit string of virtual column xv in row i, row-major taps, 1 where centre > tap (AD-Census.h:191-268)
template <int mode>
__device__ __forceinline__ uint64_t census_bits(const uint8_t* __restrict__ img, int H, int W, int i, int xv) {
    int cc = xv < 0 ? 0 : (xv > W - 1 ? W - 1 : xv);
    const int centre = mode == CM_MASK ? 0 : img[(size_t)i * W + cc];
    uint64_t v = 0;
#pragma unroll
    for (int r = -4; r <= 4; ++r) {
        const int y = i + r;
        const bool yok = (y >= 0 && y < H);
        const uint8_t* row = img + (size_t)(yok ? y : i) * W;
#pragma unroll
        for (int c = -3; c <= 3; ++c) {
            v <<= 1;
            int x = xv + c;
            bool bit;
            if (mode == CM_MASK) {
                bit = yok && x >= 0 && x < W;
            } else if (mode == CM_OWN) {
                bit = yok && x >= 0 && x < W && centre > (int)row[x < 0 ? 0 : (x >= W ? W - 1 : x)];
            } else if (mode == CM_RIGHT_EXT) {
                // x >= W never survives the anchor mask (x = j+c-d <= j+c < W); clamp only for memory safety
                int xc = x < 0 ? 0 : (x >= W ? W - 1 : x);
                bit = yok && centre > (int)row[xc];
            } else {  // CM_LEFT_EXT: overflow on the right reads column 0 (AD-Census.h:242-243)
                int xc = (x >= W || x < 0) ? 0 : x;
                bit = yok && centre > (int)row[xc];
            }
            v |= bit ? 1ull : 0ull;
        }
    }
    return v;
}

// One thread per (row, column slot) of one map.  The two "seen from the other image" maps differ from the image's own map
// only where a tap leaves the row (columns within 3 of a border, and the slots beyond it): everywhere else they copy it.
// The validity mask is all ones away from the borders.
template <int mode>
__global__ void __launch_bounds__(256) census_map_kernel(const uint8_t* __restrict__ img, int H, int W, int ncol,
                                                         const uint64_t* __restrict__ own, uint64_t* __restrict__ out) {
    int xi = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (xi >= ncol) return;
    const int xv = (mode == CM_RIGHT_EXT) ? xi - 3 : xi;  // virtual column
    const bool interior = xv >= 3 && xv <= W - 4;
    uint64_t v;
    if ((mode == CM_RIGHT_EXT || mode == CM_LEFT_EXT) && interior) v = own[(size_t)i * W + xv];
    else if (mode == CM_MASK && interior && i >= 4 && i <= H - 5) v = 0x7fffffffffffffffull;
    else v = census_bits<mode>(img, H, W, i, xv);
    out[(size_t)i * ncol + xi] = v;
}

smt_status_t k_census_maps(smt_ctx* ctx, const uint8_t* L, const uint8_t* R, int H, int W, const CensusMaps& m) {
    StageTimer t(ctx, ST_CENSUS);
    dim3 blk(128);
    SMT_LAUNCH(ctx, census_map_kernel<CM_OWN>, dim3(cdiv(W, 128), H), blk, 0, L, H, W, W, nullptr, m.CL);
    SMT_LAUNCH(ctx, census_map_kernel<CM_OWN>, dim3(cdiv(W, 128), H), blk, 0, R, H, W, W, nullptr, m.CR);
    SMT_LAUNCH(ctx, census_map_kernel<CM_MASK>, dim3(cdiv(W, 128), H), blk, 0, L, H, W, W, nullptr, m.M);
    SMT_LAUNCH(ctx, census_map_kernel<CM_RIGHT_EXT>, dim3(cdiv(W + 3, 128), H), blk, 0, R, H, W, W + 3, m.CR, m.CRx);
    SMT_LAUNCH(ctx, census_map_kernel<CM_LEFT_EXT>, dim3(cdiv(W + 4, 128), H), blk, 0, L, H, W, W + 4, m.CL, m.CLx);
    return SMT_OK;
}

smt_status_t k_ensure_lut(smt_ctx* ctx, float sigmaC, float sigmaS) {
    if (ctx->lut_sigmaC == sigmaC && ctx->lut_sigmaS == sigmaS) return SMT_OK;
    float h[320];
    // AD-Census.h:287-288 -- `1 - exp(-(AD / _sigmaC))`, every operation float32 (std::exp(float) == expf)
    for (int a = 0; a < 256; ++a) h[a] = 1 - std::exp(-((float)a / sigmaC));
    for (int c = 0; c < 64; ++c) h[256 + c] = 1 - std::exp(-((float)c / sigmaS));
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_lut, h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // h is a stack array
    ctx->lut_sigmaC = sigmaC;
    ctx->lut_sigmaS = sigmaS;
    return SMT_OK;
}

// Shared-memory slot of staged element q.  Lanes of a warp read q = base - 4*lane - k (uint64), i.e. a 32-byte
// stride: 4-way bank conflicts.  Skewing by q>>2 turns the stride into 5 eight-byte words, coprime with 16.
__device__ __forceinline__ int skew(int q) { return q + (q >> 2); }

#ifndef SMT_COST_TJ
#define SMT_COST_TJ 128  // pixels per block: the D+TJ staged strings are amortised over TJ*D cells (32: 1.21 ms, 128: 1.11 ms per frame)
#endif
constexpr int COST_TJ = SMT_COST_TJ;

// grid (ceil(W/TJ), H), 256 threads.  Each block stages the TJ own strings/masks/values and the D+TJ+2 "other image"
// strings/values one row needs, then streams TJ*D cost cells with float4 stores (VEC=4) -- D innermost, coalesced.
// PARTS: also write the AD and Hamming volumes (the int -> float conversions they need run at a quarter of the FP32 rate,
// so the cost-only instantiation does not even compute them).
template <int VEC, bool PARTS, int VIEW>
__global__ void __launch_bounds__(256)
    adcensus_cost_kernel(const uint8_t* __restrict__ own_img, const uint8_t* __restrict__ oth_img,
                         const uint64_t* __restrict__ Cown, const uint64_t* __restrict__ Cext,
                         const uint64_t* __restrict__ Mask, const float* __restrict__ lut, int H, int W, int D,
                         int nmax, float* __restrict__ cost, float* __restrict__ ad, float* __restrict__ census) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_ext = reinterpret_cast<uint64_t*>(smem_raw);               // skew(nmax)+1
    uint64_t* s_own = s_ext + skew(nmax) + 1;                              // TJ
    uint64_t* s_msk = s_own + COST_TJ;                                     // TJ
    float* s_lut = reinterpret_cast<float*>(s_msk + COST_TJ);              // 320
    uint8_t* s_oth = reinterpret_cast<uint8_t*>(s_lut + 320);              // nmax
    uint8_t* s_val = s_oth + nmax;                                         // TJ

    const int i = blockIdx.y;
    const int j0 = blockIdx.x * COST_TJ;
    const int jend = min(j0 + COST_TJ, W);
    // the strings / values of the other image at x' = max(j-d, -3) (view 0) or min(j+d, W+3) (view 1): staged for the whole
    // unclamped range j-d (j+d) of the tile with the clamp applied here, so that the cell loop indexes without one
    int xlo, ext_stride, ext_off;
    if (VIEW == 0) {
        xlo = j0 - (D - 1);
        ext_stride = W + 3;
        ext_off = 3;
    } else {
        xlo = j0;
        ext_stride = W + 4;
        ext_off = 0;
    }
    const int n = (jend - j0) + D - 1;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int xv = VIEW == 0 ? max(xlo + q, -3) : min(xlo + q, W + 3);
        s_ext[skew(q)] = Cext[(size_t)i * ext_stride + xv + ext_off];
        const int xc = xv < 0 ? 0 : (xv > W - 1 ? W - 1 : xv);
        s_oth[q] = oth_img[(size_t)i * W + xc];
    }
    for (int q = threadIdx.x; q < jend - j0; q += blockDim.x) {
        s_own[q] = Cown[(size_t)i * W + j0 + q];
        s_msk[q] = Mask[(size_t)i * W + j0 + q];
        s_val[q] = own_img[(size_t)i * W + j0 + q];
    }
    for (int q = threadIdx.x; q < 320; q += blockDim.x) s_lut[q] = lut[q];
    __syncthreads();

    const int groups = (D + VEC - 1) / VEC;  // VEC==4 requires D%4==0
    const int npj = jend - j0;
    // item = (pixel pj, group dg of VEC disparities), items of a block walked 256 at a time: (pj, dg) advance by a fixed
    // step with one carry instead of a division per item
    int pj = threadIdx.x / groups, dg = threadIdx.x - pj * groups;
    const int step_pj = 256 / groups, step_dg = 256 - step_pj * groups;
    float* cost_row = cost + (size_t)i * W * D;
    float* ad_row = PARTS && ad ? ad + (size_t)i * W * D : nullptr;
    float* census_row = PARTS && census ? census + (size_t)i * W * D : nullptr;
    for (; pj < npj; pj += step_pj, dg += step_dg) {
        if (dg >= groups) {
            dg -= groups;
            if (++pj >= npj) break;
        }
        const int j = j0 + pj;
        const uint64_t c0 = s_own[pj];
        const uint64_t m = s_msk[pj];
        const unsigned v0 = s_val[pj];
        float oc[VEC], oa[VEC], oh[VEC];
        const int q0 = VIEW == 0 ? pj + (D - 1) - dg * VEC : pj + dg * VEC;  // staged index of the group's first disparity
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int q = VIEW == 0 ? q0 - k : q0 + k;
            const uint64_t e = s_ext[skew(q)];
            const int a = (int)__usad(v0, (unsigned)s_oth[q], 0u);
            const int h = __popcll((c0 ^ e) & m);
            if (PARTS) {
                oa[k] = (float)a;
                oh[k] = (float)h;
            }
            oc[k] = __fadd_rn(s_lut[a], s_lut[256 + h]);
        }
        const int o = j * D + dg * VEC;  // inside one row: < 2^31 elements (checked by the launcher)
        if (VEC == 4) {
            __stcs(reinterpret_cast<float4*>(cost_row + o), make_float4(oc[0], oc[1], oc[2], oc[3]));
            if (PARTS && ad) __stcs(reinterpret_cast<float4*>(ad_row + o), make_float4(oa[0], oa[1], oa[2], oa[3]));
            if (PARTS && census) __stcs(reinterpret_cast<float4*>(census_row + o), make_float4(oh[0], oh[1], oh[2], oh[3]));
        } else {
            cost_row[o] = oc[0];
            if (PARTS && ad) ad_row[o] = oa[0];
            if (PARTS && census) census_row[o] = oh[0];
        }
    }
}

smt_status_t k_adcensus_cost(smt_ctx* ctx, const uint8_t* L, const uint8_t* R, const CensusMaps& m, int H, int W, int D,
                             int view, float* cost, float* ad, float* census) {
    StageTimer t(ctx, ST_COST);
    if ((long long)W * D >= 0x7fffffffLL) return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "adcensus_cost: a row of %d x %d cells", W, D);
    const int nmax = D + COST_TJ + 3;
    size_t smem = sizeof(uint64_t) * (size_t)(nmax + (nmax >> 2) + 1 + 2 * COST_TJ) + 320 * sizeof(float) + nmax + COST_TJ;
    smem = (smem + 15) & ~size_t(15);
    const uint8_t* own = view == 0 ? L : R;
    const uint8_t* oth = view == 0 ? R : L;
    const uint64_t* Cown = view == 0 ? m.CL : m.CR;
    const uint64_t* Cext = view == 0 ? m.CRx : m.CLx;
    dim3 grid(cdiv(W, COST_TJ), H);
    const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0) &&
                      (!ad || (reinterpret_cast<uintptr_t>(ad) & 15) == 0) &&
                      (!census || (reinterpret_cast<uintptr_t>(census) & 15) == 0);
    const bool parts = ad != nullptr || census != nullptr;
#define SMT_COST_LAUNCH2(VEC_, PARTS_, VIEW_)                                                                                                \
    do {                                                                                                                             \
        if (smem > 48 * 1024)                                                                                                        \
            SMT_CUDA(ctx, cudaFuncSetAttribute(adcensus_cost_kernel<VEC_, PARTS_, VIEW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        SMT_LAUNCH(ctx, (adcensus_cost_kernel<VEC_, PARTS_, VIEW_>), grid, 256, smem, own, oth, Cown, Cext, m.M, ctx->d_lut, H, W, D, nmax, \
                   cost, ad, census);                                                                                                \
    } while (0)
#define SMT_COST_VIEW(VEC_, PARTS_)                \
    do {                                           \
        if (view == 0) SMT_COST_LAUNCH2(VEC_, PARTS_, 0); \
        else SMT_COST_LAUNCH2(VEC_, PARTS_, 1);    \
    } while (0)
    if (vec4 && parts) SMT_COST_VIEW(4, true);
    else if (vec4) SMT_COST_VIEW(4, false);
    else if (parts) SMT_COST_VIEW(1, true);
    else SMT_COST_VIEW(1, false);
#undef SMT_COST_VIEW
#undef SMT_COST_LAUNCH2
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- WTA
// One warp per pixel; lane l scans d = l, l+32, ... (coalesced 128 B per step) keeping its first minimum, then the
// warp reduces (value, d) lexicographically, which is the reference's strict-'>' first minimum (AD-Census.h:359-372).
#ifndef SMT_WTA_NP
#define SMT_WTA_NP 4
#endif
constexpr int WTA_NP = SMT_WTA_NP;  // pixels per warp: their loads are issued together (the kernel is bound by load latency)
template <bool VEC4>
__global__ void __launch_bounds__(256) wta_first_min_kernel(const float* __restrict__ vol, size_t npix, int D,
                                                            float* __restrict__ disp) {
    const size_t pix0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * WTA_NP;
    if (pix0 >= npix) return;
    const int lane = threadIdx.x & 31;
    float best[WTA_NP];
    int bd[WTA_NP];
#pragma unroll
    for (int k = 0; k < WTA_NP; ++k) {
        best[k] = INFINITY;
        bd[k] = 0x7fffffff;
    }
    if (VEC4) {
        // lane l owns d = 4l .. 4l+3 (+128, ...): 512 bytes per warp and request, ascending d inside the lane
        for (int d = lane * 4; d < D; d += 128) {
            float4 v[WTA_NP];
#pragma unroll
            for (int k = 0; k < WTA_NP; ++k)
                v[k] = __ldcs(reinterpret_cast<const float4*>(vol + min(pix0 + k, npix - 1) * D + d));
#pragma unroll
            for (int k = 0; k < WTA_NP; ++k) {
                const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (e[q] < best[k] || bd[k] == 0x7fffffff) {
                        best[k] = e[q];
                        bd[k] = d + q;
                    }
            }
        }
    } else {
        for (int d = lane; d < D; d += 32) {
            float v[WTA_NP];
#pragma unroll
            for (int k = 0; k < WTA_NP; ++k) v[k] = __ldcs(vol + min(pix0 + k, npix - 1) * D + d);
#pragma unroll
            for (int k = 0; k < WTA_NP; ++k)
                if (v[k] < best[k] || bd[k] == 0x7fffffff) {
                    best[k] = v[k];
                    bd[k] = d;
                }
        }
    }
#pragma unroll
    for (int k = 0; k < WTA_NP; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best[k], o);
            int od = __shfl_xor_sync(0xffffffffu, bd[k], o);
            if (od != 0x7fffffff && (bd[k] == 0x7fffffff || ov < best[k] || (ov == best[k] && od < bd[k]))) {
                best[k] = ov;
                bd[k] = od;
            }
        }
        if (lane == 0 && pix0 + k < npix) disp[pix0 + k] = (float)bd[k];
    }
}

smt_status_t k_wta(smt_ctx* ctx, const float* vol, int H, int W, int D, float* disp) {
    StageTimer t(ctx, ST_WTA);
    size_t npix = (size_t)H * W;
    const unsigned nblk = (unsigned)cdiv((long long)npix, 8 * WTA_NP);
    if (D % 4 == 0 && (reinterpret_cast<uintptr_t>(vol) & 15) == 0)
        SMT_LAUNCH(ctx, wta_first_min_kernel<true>, nblk, 256, 0, vol, npix, D, disp);
    else
        SMT_LAUNCH(ctx, wta_first_min_kernel<false>, nblk, 256, 0, vol, npix, D, disp);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

smt_status_t smt_adcensus_cost(smt_ctx* ctx, const float* left, const float* right, int32_t rows, int32_t cols, int32_t D,
                               float sigmaC, float sigmaS, int32_t view, float* cost, float* ad, float* census,
                               smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, left && right && cost, SMT_ERR_INVALID_ARG, "smt_adcensus_cost: null image or cost pointer");
    SMT_REQUIRE(ctx, rows > 0 && cols > 0 && D > 0 && (view == 0 || view == 1), SMT_ERR_INVALID_ARG,
                "smt_adcensus_cost: rows, cols, D must be positive and view 0 or 1");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols, nvol = npix * D;
    Marshal ms(ctx, space);
    const float* dl = ms.in(left, npix);
    const float* dr = ms.in(right, npix);
    float* dcost = ms.out(cost, nvol);
    float* dad = ms.out(ad, nvol);
    float* dcen = ms.out(census, nvol);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, u8, 2 * npix);
    SMT_ALLOC(ctx, maps, census_maps_bytes(rows, cols));
    uint8_t* L8 = u8.as<uint8_t>();
    uint8_t* R8 = L8 + npix;
    SMT_TRY(k_f32_to_u8(ctx, dl, npix, L8));
    SMT_TRY(k_f32_to_u8(ctx, dr, npix, R8));
    SMT_TRY(k_ensure_lut(ctx, sigmaC, sigmaS));
    CensusMaps cm;
    census_maps_carve(maps.p, rows, cols, &cm);
    SMT_TRY(k_census_maps(ctx, L8, R8, rows, cols, cm));
    SMT_TRY(k_adcensus_cost(ctx, L8, R8, cm, rows, cols, D, view, dcost, dad, dcen));
    return ms.finish();
}

smt_status_t smt_census9x7_u64(smt_ctx* ctx, const float* img, int32_t rows, int32_t cols, uint64_t* out,
                               smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, img && out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_census9x7_u64: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* di = ms.in(img, npix);
    uint64_t* dout = ms.out(out, npix);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, u8, npix);
    SMT_TRY(k_f32_to_u8(ctx, di, npix, u8.as<uint8_t>()));
    {
        StageTimer t(ctx, ST_CENSUS);
        SMT_LAUNCH(ctx, census_map_kernel<CM_OWN>, dim3(cdiv(cols, 128), rows), 128, 0, u8.as<uint8_t>(), rows, cols, cols, (const uint64_t*)nullptr, dout);
    }
    return ms.finish();
}

smt_status_t smt_wta_first_min(smt_ctx* ctx, const float* vol, int32_t rows, int32_t cols, int32_t D, float* disp,
                               smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, vol && disp && rows > 0 && cols > 0 && D > 0, SMT_ERR_INVALID_ARG, "smt_wta_first_min: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* dv = ms.in(vol, npix * D);
    float* dd = ms.out(disp, npix);
    SMT_TRY(ms.check());
    SMT_TRY(k_wta(ctx, dv, rows, cols, D, dd));
    return ms.finish();
}

}  // extern "C"
