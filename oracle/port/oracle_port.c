/* ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's hot-path algorithms; see oracle_port.h for the parity status.
 * Each function cites the reference file:line it restates.  Floating-point code keeps the reference's
 * operation order and types (build with -ffp-contract=off, no -ffast-math) so results are bit-identical.
 */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int orc_abi_version(void) { return 1; }
int orc_num_threads(void) { return 1; }

#define VOL(i, j, d) ((((size_t)(i)) * W + (size_t)(j)) * D + (size_t)(d))

/* ------------------------------------------------------------------------------------------------ AD-Census
 * AD-CensusV1/AD-Census.h.  Images are float arrays holding grey levels. */

/* ComputeAD (:75-101) / ComputeADRight (:103-129): an out-of-range match copies the previous disparity's cost,
 * which is the clamped match. */
static void ad_volume(const float* own, const float* oth, int H, int W, int D, int sign, float* out) {
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const float v = own[(size_t)i * W + j];
            for (int d = 0; d < D; ++d) {
                const int x = j + sign * d;
                if (x < 0 || x >= W) {
                    out[VOL(i, j, d)] = out[VOL(i, j, d - 1)]; /* :88-92, :116-120 */
                } else {
                    const float o = oth[(size_t)i * W + x];
                    /* left view: |L - R| (:96); right view: |L - R| with L the shifted one (:124) */
                    out[VOL(i, j, d)] = sign < 0 ? fabsf(v - o) : fabsf(o - v);
                }
            }
        }
}

/* ComputeCensus9x7 (:142-204, right = 0) and ComputeCensus9x7Right (:207-269, right = 1).
 * Both 63-tap strings are rebuilt for every (pixel, d); a tap is skipped (bit 0 in both strings) when it
 * leaves the image in the ANCHOR's frame (:173, :238).  Left view: the right image's centre and taps clamp to
 * column 0 (:159-164, :177-182).  Right view: the left centre clamps to column W-1 (:224-229) but an
 * overflowing tap reads column 0 (:242-243). */
static void census_volume(const float* L, const float* R, int H, int W, int D, int right, float* out) {
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const float own_c = right ? R[(size_t)i * W + j] : L[(size_t)i * W + j];
            for (int d = 0; d < D; ++d) {
                float oth_c;
                if (!right)
                    oth_c = (j - d < 0) ? R[(size_t)i * W] : R[(size_t)i * W + j - d];
                else
                    oth_c = (j + d >= W) ? L[(size_t)i * W + W - 1] : L[(size_t)i * W + j + d];
                uint64_t a = 0, b = 0;
                for (int r = -4; r <= 4; ++r)
                    for (int c = -3; c <= 3; ++c) {
                        a <<= 1;
                        b <<= 1;
                        const int y = i + r, x = j + c;
                        if (y < 0 || y >= H || x < 0 || x >= W) continue;
                        float own_t, oth_t;
                        if (!right) {
                            own_t = L[(size_t)y * W + x];
                            oth_t = (x - d < 0) ? R[(size_t)y * W] : R[(size_t)y * W + x - d];
                        } else {
                            own_t = R[(size_t)y * W + x];
                            oth_t = (x + d >= W) ? L[(size_t)y * W] : L[(size_t)y * W + x + d];
                        }
                        if (own_c > own_t) a += 1;
                        if (oth_c > oth_t) b += 1;
                    }
                uint64_t x = a ^ b;
                int n = 0;
                while (x) { /* :194-198 */
                    ++n;
                    x &= x - 1;
                }
                out[VOL(i, j, d)] = (float)n;
            }
        }
}

/* first-minimum WTA with a strict comparison (:346-380; CrossArm.cpp:33-57; ScanlineOptimizer.h:40-64) */
void orc_wta(const float* vol, int H, int W, int D, float* disp) {
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float best = vol[VOL(i, j, 0)];
            float bd = 0;
            for (int d = 1; d < D; ++d) {
                const float v = vol[VOL(i, j, d)];
                if (best > v) {
                    best = v;
                    bd = (float)d;
                }
            }
            disp[(size_t)i * W + j] = bd;
        }
}

void orc_adcensus(const float* left, const float* right, int H, int W, int D, float sigmaC, float sigmaS, int do_left,
                  int do_right, float* costL, float* costR, float* adL, float* adR, float* censusL, float* censusR,
                  float* wtaL, float* wtaR) {
    const size_t n = (size_t)H * W * D;
    float* out_cost[2] = {costL, costR};
    float* out_ad[2] = {adL, adR};
    float* out_cen[2] = {censusL, censusR};
    float* out_wta[2] = {wtaL, wtaR};
    const int doit[2] = {do_left, do_right};
    for (int v = 0; v < 2; ++v) {
        if (!doit[v]) {
            /* Initialize zero-fills every volume (:334-342); WTA of an all-zero volume is 0 */
            if (out_cost[v]) memset(out_cost[v], 0, n * sizeof(float));
            if (out_ad[v]) memset(out_ad[v], 0, n * sizeof(float));
            if (out_cen[v]) memset(out_cen[v], 0, n * sizeof(float));
            if (wtaL && wtaR) memset(out_wta[v], 0, (size_t)H * W * sizeof(float));
            continue;
        }
        float* ad = (float*)malloc(n * sizeof(float));
        float* cen = (float*)malloc(n * sizeof(float));
        float* cost = (float*)malloc(n * sizeof(float));
        if (v == 0)
            ad_volume(left, right, H, W, D, -1, ad);
        else
            ad_volume(right, left, H, W, D, +1, ad);
        census_volume(left, right, H, W, D, v, cen);
        for (size_t k = 0; k < n; ++k) {
            /* ComputeADcensus (:287-289): every operation in float32 (std::exp(float)) */
            const float a = 1 - expf(-(ad[k] / sigmaC));
            const float c = 1 - expf(-(cen[k] / sigmaS));
            cost[k] = a + c;
        }
        if (wtaL && wtaR) orc_wta(cost, H, W, D, out_wta[v]);
        if (out_cost[v]) memcpy(out_cost[v], cost, n * sizeof(float));
        if (out_ad[v]) memcpy(out_ad[v], ad, n * sizeof(float));
        if (out_cen[v]) memcpy(out_cen[v], cen, n * sizeof(float));
        free(ad);
        free(cen);
        free(cost);
    }
}

/* ------------------------------------------------------------------------------------------------ CrossAggregator
 * CBLSM/cross_aggregator.cpp */
static int cdist(const uint8_t* a, const uint8_t* b) { /* cross_aggregator.h:78-80 */
    int m = abs((int)a[0] - (int)b[0]);
    const int g = abs((int)a[1] - (int)b[1]), r = abs((int)a[2] - (int)b[2]);
    if (g > m) m = g;
    if (r > m) m = r;
    return m;
}

/* FindHorizontalArm / FindVerticalArm (:135-269): step = +-1 pixel or +-W pixels, room = pixels before the border */
static uint8_t grow_arm(const uint8_t* img, size_t p0, long step, int room, int L1, int L2, int t1, int t2) {
    const int lim = L1 < 255 ? L1 : 255; /* min(cross_L1_, MAX_ARM_LENGTH) (:151) */
    const uint8_t* c0 = img + 3 * p0;
    const uint8_t* last = c0;
    int arm = 0;
    for (int n = 0; n < lim; ++n) {
        if (n >= room) break; /* :154-163 */
        const uint8_t* c = img + 3 * (size_t)((long)p0 + step * (n + 1));
        const int d1 = cdist(c, c0);
        if (d1 >= t1) break;                    /* :169-172 */
        if (n > 0 && cdist(c, last) >= t1) break; /* :175-180 */
        if (n + 1 > L2 && d1 >= t2) break;      /* :183-187 */
        ++arm;
        last = c;
    }
    return (uint8_t)arm;
}

int orc_cross_aggregate(const uint8_t* bgrL, const uint8_t* bgrR, const float* cost_init, int W, int H, int D, int L1,
                        int L2, int t1, int t2, int iters, float* cost_out, uint8_t* arms_out, uint16_t* sup0_out,
                        uint16_t* sup1_out) {
    (void)bgrR; /* only img_left_ is read (:138,:206) */
    const size_t npix = (size_t)W * H;
    if ((long)W * H <= 0 || D <= 0) return 0; /* Initialize (:26-31) */
    uint8_t* arms = (uint8_t*)malloc(npix * 4);
    uint16_t* sup[2] = {(uint16_t*)malloc(npix * 2), (uint16_t*)malloc(npix * 2)};
    uint16_t* suptmp = (uint16_t*)malloc(npix * 2);
    float* t0 = (float*)malloc(npix * sizeof(float));
    float* t1v = (float*)malloc(npix * sizeof(float));
    float* agg = (float*)malloc(npix * D * sizeof(float));
    /* BuildArms (:76-86) */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t p = (size_t)y * W + x;
            arms[4 * p + 0] = grow_arm(bgrL, p, -1, x, L1, L2, t1, t2);
            arms[4 * p + 1] = grow_arm(bgrL, p, +1, W - 1 - x, L1, L2, t1, t2);
            arms[4 * p + 2] = grow_arm(bgrL, p, -(long)W, y, L1, L2, t1, t2);
            arms[4 * p + 3] = grow_arm(bgrL, p, +(long)W, H - 1 - y, L1, L2, t1, t2);
        }
    /* ComputeSupPixelCount (:271-325): id 0 = horizontal first, id 1 = vertical first */
    for (int id = 0; id < 2; ++id)
        for (int pass = 0; pass < 2; ++pass)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const size_t p = (size_t)y * W + x;
                    const int l = arms[4 * p], r = arms[4 * p + 1], t = arms[4 * p + 2], b = arms[4 * p + 3];
                    const int horizontal = (id == 0) == (pass == 0);
                    int count = 0;
                    if (pass == 0) {
                        count = horizontal ? (l + r + 1) : (t + b + 1);
                        suptmp[p] = (uint16_t)count;
                    } else {
                        if (horizontal)
                            for (int s = -l; s <= r; ++s) count += suptmp[p + s];
                        else
                            for (int s = -t; s <= b; ++s) count += suptmp[(size_t)(y + s) * W + x];
                        sup[id][p] = (uint16_t)count;
                    }
                }
    /* Aggregate (:89-118) */
    memcpy(agg, cost_init, npix * D * sizeof(float));
    int hfirst = 1;
    for (int k = 0; k < iters; ++k) {
        for (int d = 0; d < D; ++d) {
            /* AggregateInArms (:327-394) */
            for (size_t p = 0; p < npix; ++p) t0[p] = agg[p * D + d];
            for (int pass = 0; pass < 2; ++pass)
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < W; ++x) {
                        const size_t p = (size_t)y * W + x;
                        const int horizontal = (hfirst != 0) == (pass == 0);
                        const float* src = pass == 0 ? t0 : t1v;
                        float acc = 0.0f;
                        if (horizontal)
                            for (int s = -(int)arms[4 * p]; s <= (int)arms[4 * p + 1]; ++s) acc += src[p + s];
                        else
                            for (int s = -(int)arms[4 * p + 2]; s <= (int)arms[4 * p + 3]; ++s)
                                acc += src[(size_t)(y + s) * W + x];
                        if (pass == 0)
                            t1v[p] = acc;
                        else
                            agg[p * D + d] = acc / sup[hfirst ? 0 : 1][p]; /* float / int(uint16) (:389) */
                    }
        }
        hfirst = !hfirst;
    }
    if (cost_out) memcpy(cost_out, agg, npix * D * sizeof(float));
    if (arms_out) memcpy(arms_out, arms, npix * 4);
    if (sup0_out) memcpy(sup0_out, sup[0], npix * 2);
    if (sup1_out) memcpy(sup1_out, sup[1], npix * 2);
    free(arms);
    free(sup[0]);
    free(sup[1]);
    free(suptmp);
    free(t0);
    free(t1v);
    free(agg);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ ScanlineOptimizer
 * AD-CensusV1/ScanlineOptimizer.h */
static float fmin2(float a, float b) { return b < a ? b : a; } /* std::min */

/* One path: n pixels starting at pixel p0 with pixel stride pstep; grey values start at g0 with stride gstep.
 * vertical != 0 selects the up/down variant: reference grey fixed at the start pixel (:210,:232), l2 = prev[d]+p1 (:238). */
static void scan_one_path(const float* cost, const float* gray, int D, long p0, long pstep, long g0, long gstep, int n,
                          float p1, float p2init, int vertical, float* out, float* prev /* D+2 */) {
    for (int d = 0; d < D + 2; ++d) prev[d] = 65535.f; /* 0xffff (:151,:213) */
    memcpy(out + (size_t)p0 * D, cost + (size_t)p0 * D, (size_t)D * sizeof(float)); /* :153,:215 */
    memcpy(prev + 1, out + (size_t)p0 * D, (size_t)D * sizeof(float));
    float m = 65535.f;
    for (int d = 0; d < D + 2; ++d) m = fmin2(prev[d], m); /* :162-166 */
    float glast = gray[g0];
    for (int s = 1; s < n; ++s) {
        const long p = p0 + pstep * s;
        const float g = gray[g0 + gstep * s];
        float q = p2init / (fabsf(g - glast) + 1); /* :171,:232 */
        const float p2 = p1 > q ? p1 : q;          /* std::max(p1, ...) */
        if (!vertical) glast = g;                  /* :172 (never updated in up/down) */
        float mc = 65535.f;
        float* o = out + (size_t)p * D;
        const float* c = cost + (size_t)p * D;
        for (int d = 0; d < D; ++d) {
            const float l1 = prev[d + 1];
            const float l2 = (vertical ? prev[d + 1] : prev[d]) + p1; /* :238 vs :177 */
            const float l3 = prev[d + 2] + p1;
            const float l4 = m + p2;
            const float v = c[d] + fmin2(fmin2(l1, l2), fmin2(l3, l4)) - m; /* :180,:241 */
            o[d] = v;
            mc = fmin2(mc, v);
        }
        m = mc;
        memcpy(prev + 1, o, (size_t)D * sizeof(float));
    }
}

void orc_scanline(const float* cost, const float* gray, int H, int W, int D, int p1i, int p2i, float* left,
                  float* right, float* up, float* down, float* sum, float* disp) {
    if (H <= 0 || W <= 0 || D <= 0) return;
    const size_t n = (size_t)H * W * D;
    const float p1 = (float)p1i, p2 = (float)p2i; /* int members (:66,:132-133) */
    float* vol[4];
    float* user[4] = {left, right, up, down};
    for (int k = 0; k < 4; ++k) vol[k] = (float*)calloc(n, sizeof(float));
    float* prev = (float*)malloc((size_t)(D + 2) * sizeof(float));
    for (int i = 0; i < H; ++i) { /* ScanLineLeftRight (:130-192) */
        scan_one_path(cost, gray, D, (long)i * W, +1, (long)i * W, +1, W, p1, p2, 0, vol[0], prev);
        scan_one_path(cost, gray, D, (long)i * W + W - 1, -1, (long)i * W + W - 1, -1, W, p1, p2, 0, vol[1], prev);
    }
    for (int j = 0; j < W; ++j) { /* ScanLineUpDown (:194-253): grey pointer moves by +-1 element (:221,:250) */
        scan_one_path(cost, gray, D, j, +W, j, +1, H, p1, p2, 1, vol[2], prev);
        scan_one_path(cost, gray, D, (long)(H - 1) * W + j, -W, (long)(H - 1) * W + j, -1, H, p1, p2, 1, vol[3], prev);
    }
    float* s = (float*)malloc(n * sizeof(float));
    for (size_t k = 0; k < n; ++k) s[k] = vol[0][k] + vol[1][k] + vol[2][k] + vol[3][k]; /* :124 */
    if (disp) orc_wta(s, H, W, D, disp);
    if (sum) memcpy(sum, s, n * sizeof(float));
    for (int k = 0; k < 4; ++k) {
        if (user[k]) memcpy(user[k], vol[k], n * sizeof(float));
        free(vol[k]);
    }
    free(s);
    free(prev);
}

/* ------------------------------------------------------------------------------------------------ PostProcessing.h */
static void push_pair(int32_t* list, int32_t* n, int i, int j) {
    if (list) {
        list[2 * (size_t)*n] = i;
        list[2 * (size_t)*n + 1] = j;
    }
    ++*n;
}

/* LeftRightConsistency (:72-135): in place, row-major, +inf marks invalid */
void orc_lrcheck(float* dispL, const float* dispR, int W, int H, int gate, int32_t* occ, int32_t* n_occ, int32_t* mis,
                 int32_t* n_mis) {
    int32_t no = 0, nm = 0;
    const float thr = (float)gate;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float* d = &dispL[(size_t)i * W + j];
            if (*d == INFINITY) { /* :90-93 */
                push_pair(mis, &nm, i, j);
                continue;
            }
            const int cr = (int)(j - *d + 0.5); /* :96 */
            if (cr >= 0 && cr < W) {
                const float dr = dispR[(size_t)i * W + cr];
                if (fabsf(*d - dr) > thr) { /* :103 */
                    const int crl = (int)(cr + dr + 0.5); /* :110 */
                    if (crl > 0 && crl < W) {
                        if (dispL[(size_t)i * W + crl] > *d) /* reads the array being modified (:112) */
                            push_pair(occ, &no, i, j);
                        else
                            push_pair(mis, &nm, i, j);
                    } else {
                        push_pair(mis, &nm, i, j);
                    }
                    *d = INFINITY;
                }
            } else {
                *d = INFINITY; /* :130 */
                push_pair(mis, &nm, i, j);
            }
        }
    if (n_occ) *n_occ = no;
    if (n_mis) *n_mis = nm;
}

/* LeftAndRightConsistency (:10-70): separate output, >= gate, no +inf test */
void orc_lrcheck_and(const float* dispL, const float* dispR, float* last, int W, int H, float gate, int32_t* occ,
                     int32_t* n_occ, int32_t* mis, int32_t* n_mis) {
    int32_t no = 0, nm = 0;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const float d = dispL[(size_t)i * W + j];
            float* lv = &last[(size_t)i * W + j];
            const int cr = (int)(j - d + 0.5); /* :25 */
            if (cr >= 0 && cr < W) {
                const float dr = dispR[(size_t)i * W + cr];
                if (fabsf(d - dr) >= gate) { /* :33 */
                    const int crl = (int)(cr + dr + 0.5);
                    if (crl > 0 && crl < W) {
                        if (dispL[(size_t)i * W + crl] > d)
                            push_pair(occ, &no, i, j);
                        else
                            push_pair(mis, &nm, i, j);
                    } else {
                        push_pair(mis, &nm, i, j);
                    }
                    *lv = 0;
                } else {
                    *lv = d;
                }
            } else {
                *lv = 0;
                push_pair(mis, &nm, i, j);
            }
        }
    if (n_occ) *n_occ = no;
    if (n_mis) *n_mis = nm;
}

/* RemoveSpeckles (:250-311): breadth-first region growing over the 8-neighbourhood */
void orc_remove_speckles(float* disp, int W, int H, int diff, unsigned min_area, int invalid_val) {
    if (W <= 0 || H <= 0) return;
    const size_t n = (size_t)W * H;
    unsigned char* seen = (unsigned char*)calloc(n, 1);
    int32_t* queue = (int32_t*)malloc(n * sizeof(int32_t));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const size_t p = (size_t)i * W + j;
            if (seen[p] || disp[p] == invalid_val) continue; /* float == int (:262) */
            size_t head = 0, tail = 0;
            queue[tail++] = (int32_t)p;
            seen[p] = 1;
            while (head < tail) {
                const int32_t q = queue[head++];
                const int r0 = q / W, c0 = q % W;
                const float base = disp[q];
                for (int r = -1; r <= 1; ++r)
                    for (int c = -1; c <= 1; ++c) {
                        if (r == 0 && c == 0) continue;
                        const int rr = r0 + r, cc = c0 + c;
                        if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
                        const size_t t = (size_t)rr * W + cc;
                        if (!seen[t] && disp[t] != invalid_val && fabsf(disp[t] - base) <= diff) { /* :290-292 */
                            queue[tail++] = (int32_t)t;
                            seen[t] = 1;
                        }
                    }
            }
            if (tail < min_area) /* :304 */
                for (size_t k = 0; k < tail; ++k) disp[queue[k]] = (float)invalid_val;
        }
    free(seen);
    free(queue);
}

/* The shipped call `RemoveSpeckles(disp, W, H, 1, 30, Invalid_Float)` (main.cpp:93) binds +inf to `const int&`;
 * g++ folds that conversion to INT_MAX, which is what the compiled reference uses. */
void orc_remove_speckles_inf(float* disp, int W, int H, int diff, unsigned min_area) {
    orc_remove_speckles(disp, W, H, diff, min_area, 2147483647);
}

static int cmp_float(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* MedianFilter (:314-344): in-bounds taps, element [n/2] of the sorted window */
void orc_median(const float* in, float* out, int W, int H, int k) {
    const int r = k / 2;
    float* w = (float*)malloc((size_t)k * k * sizeof(float) + sizeof(float));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            int n = 0;
            for (int a = -r; a <= r; ++a)
                for (int b = -r; b <= r; ++b) {
                    const int y = i + a, x = j + b;
                    if (y >= 0 && y < H && x >= 0 && x < W) w[n++] = in[(size_t)y * W + x];
                }
            qsort(w, (size_t)n, sizeof(float), cmp_float);
            out[(size_t)i * W + j] = w[n / 2];
        }
    free(w);
}

/* ------------------------------------------------------------------------------------------------ grey/colour arms
 * AD-CensusV1/CrossArm.cpp:147-598 (CrossArmAggregation::Compute{Left,Right,Top,Button}ArmLength) and
 * CBLSM/CBLSM.h:536-966 (ArmLength{R,L,Up,Down}).  One walk per (pixel, direction):
 *   offsets k = 1, 2, ...; the threshold drops to 6 for good once k exceeds `sec` (the assignment sticks in *tao);
 *   the arm is capped at `maxlen`; the walk stops at the image border or at the first pixel whose largest channel
 *   difference exceeds the threshold; a pixel at least two away from the border keeps arm 1 when the very first
 *   neighbour already fails. */
static int walk_grey_arm(const uint8_t* img, int ch, long p0, long step, int room, int minrule, int sec, int maxlen,
                         int* tao) {
    const uint8_t* c0 = img + (size_t)p0 * ch;
    int saved = 0, k = 0;
    while (k <= room) {          /* `while (j - offset >= 0)` with offset == k */
        saved = k;
        ++k;
        if (k > sec) {           /* CrossArm.cpp:168-173 */
            *tao = 6;
            if (k > maxlen) break;
        }
        if (k <= room) {
            const uint8_t* c = img + (size_t)(p0 + step * k) * ch;
            int m = 0;
            for (int t = 0; t < ch; ++t) {
                const int dlt = abs((int)c0[t] - (int)c[t]);
                if (dlt > m) m = dlt;
            }
            if (m > *tao) {      /* :184-197 */
                if (minrule && saved < 1) saved = 1;
                break;
            }
        } else {
            break;
        }
    }
    return saved;
}

void orc_cblsm_arms(const uint8_t* img, int ch, int H, int W, int tao0, int maxLength, int secLength, int32_t* armL,
                    int32_t* armR, int32_t* armUp, int32_t* armDown) {
    /* `tao` is a by-value uchar parameter: it sticks at 6 only within one call (CBLSM.h:536, :558) */
    int32_t* outs[4] = {armL, armR, armUp, armDown};
    for (int dir = 0; dir < 4; ++dir) {
        if (!outs[dir]) continue;
        int tao = (uint8_t)tao0;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const long p = (long)i * W + j;
                int v;
                if (dir == 0) v = walk_grey_arm(img, ch, p, -1, j, j - 1 >= 1, secLength, maxLength, &tao);
                else if (dir == 1) v = walk_grey_arm(img, ch, p, +1, W - 1 - j, j + 1 < W - 1, secLength, maxLength, &tao);
                else if (dir == 2) v = walk_grey_arm(img, ch, p, -W, i, i - 1 >= 1, secLength, maxLength, &tao);
                else v = walk_grey_arm(img, ch, p, +W, H - 1 - i, i + 1 < H - 1, secLength, maxLength, &tao);
                outs[dir][p] = v;
            }
    }
}

/* CrossArmAggregation: `_tao` is a MEMBER, so the drop to 6 survives across pixels and across the four passes, which
 * main.cpp:69-72 runs in the order Left, Right, Top, Button.  The Right pass uses `col = _row` (CrossArm.cpp:265): it
 * visits columns j < H only, treats column H as the border and stores at i*H + j.  mode 1 = AggregationVertical
 * (:60-102), mode 2 = Aggregation (:104-145); both address the slice with a flat index, so a window running past the
 * end of a row continues in the next one.  Returns the final _tao. */
int orc_crossarm(const uint8_t* img, int ch, int H, int W, int D, int tao0, int32_t* armL, int32_t* armR, int32_t* armT,
                 int32_t* armB, int mode, const float* vol_in, float* vol_out) {
    const size_t npix = (size_t)H * W;
    int tao = tao0;
    int32_t* L = (int32_t*)calloc(npix, sizeof(int32_t));
    int32_t* R = (int32_t*)calloc(npix, sizeof(int32_t));
    int32_t* T = (int32_t*)calloc(npix, sizeof(int32_t));
    int32_t* B = (int32_t*)calloc(npix, sizeof(int32_t));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
            L[(size_t)i * W + j] = walk_grey_arm(img, ch, (long)i * W + j, -1, j, j - 1 >= 1, 17, 34, &tao);
    {
        const int col = H; /* the bug: col = _row */
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < col; ++j)
                R[(size_t)i * col + j] =
                    walk_grey_arm(img, ch, (long)i * W + j, +1, col - 1 - j, j + 1 < col - 1, 17, 34, &tao);
    }
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
            T[(size_t)i * W + j] = walk_grey_arm(img, ch, (long)i * W + j, -W, i, i - 1 >= 1, 17, 34, &tao);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
            B[(size_t)i * W + j] = walk_grey_arm(img, ch, (long)i * W + j, +W, H - 1 - i, i + 1 < H - 1, 17, 34, &tao);
    if (mode) {
        float* layer = (float*)malloc(npix * sizeof(float));
        for (int d = 0; d < D; ++d) {
            for (size_t p = 0; p < npix; ++p) layer[p] = vol_in[p * D + d];
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    const size_t p = (size_t)i * W + j;
                    const int l = L[p], r = R[p], up = T[p], down = B[p];
                    float value = 0;
                    int number = 0;
                    if (mode == 1) { /* column-major, inclusive bounds (:88-95) */
                        for (int a = -l; a <= r; ++a)
                            for (int b = -up; b <= down; ++b) {
                                value = value + layer[(long)(i + b) * W + j + a];
                                ++number;
                            }
                    } else { /* row-major, EXCLUSIVE upper bounds (:131-138) */
                        for (int b = -up; b < down; ++b)
                            for (int a = -l; a < r; ++a) {
                                value = value + layer[(long)(i + b) * W + j + a];
                                ++number;
                            }
                    }
                    vol_out[p * D + d] = value / number;
                }
        }
        free(layer);
    }
    if (armL) memcpy(armL, L, npix * sizeof(int32_t));
    if (armR) memcpy(armR, R, npix * sizeof(int32_t));
    if (armT) memcpy(armT, T, npix * sizeof(int32_t));
    if (armB) memcpy(armB, B, npix * sizeof(int32_t));
    free(L);
    free(R);
    free(T);
    free(B);
    return tao;
}

/* ------------------------------------------------------------------------------------------------ CBLSM.h live part */
/* ComputeAD / ComputeADRight (:327-381): integer |L - R| from uchar images, clamp-by-copy at the border */
void orc_cblsm_ad(const uint8_t* left, const uint8_t* right, int H, int W, int D, float* adL, float* adR) {
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
            for (int d = 0; d < D; ++d) {
                if (adL) adL[VOL(i, j, d)] = (j - d < 0) ? adL[VOL(i, j, d - 1)]
                                                         : (float)abs((int)left[(size_t)i * W + j] - (int)right[(size_t)i * W + j - d]);
                if (adR) adR[VOL(i, j, d)] = (j + d >= W) ? adR[VOL(i, j, d - 1)]
                                                          : (float)abs((int)left[(size_t)i * W + j + d] - (int)right[(size_t)i * W + j]);
            }
}

/* costAggregationV5 (:1179-1224): mean over the rectangle spanned by the pixel's four arms, row-major accumulation */
void orc_cblsm_aggregate_v5(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR,
                            const int32_t* armUp, const int32_t* armDown, int D, int H, int W) {
    const size_t npix = (size_t)H * W;
    float* layer = (float*)malloc(npix * sizeof(float));
    for (int d = 0; d < D; ++d) {
        for (size_t p = 0; p < npix; ++p) layer[p] = vol_in[p * D + d];
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const size_t p = (size_t)i * W + j;
                float value = 0;
                int number = 0;
                for (int b = -armUp[p]; b <= armDown[p]; ++b)
                    for (int a = -armL[p]; a <= armR[p]; ++a) {
                        value = value + layer[(long)(i + b) * W + j + a];
                        ++number;
                    }
                vol_out[p * D + d] = value / number;
            }
    }
    free(layer);
}

/* ComputeDispOringin (:383-407): first minimum, strict '<' */
void orc_cblsm_wta(const float* vol, float* disp, int D, int H, int W) { orc_wta(vol, H, W, D, disp); }

/* ------------------------------------------------------------------------------------------------ SAD (SAD/Sad.h)
 * Inputs are replicate-padded grey images [Hp][Wp], pad w = winsize+1 per side (SADmain.cpp:47-48). */
static float sad_window(const uint8_t* A, const uint8_t* B, int Wp, int ai, int aj, int bi, int bj, int k) {
    int s = 0; /* sadvalue (:15-20): cv::abs(a-b) is absdiff for 8-bit, cv::sum -> int -> float */
    for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x) s += abs((int)A[(size_t)(ai + y) * Wp + aj + x] - (int)B[(size_t)(bi + y) * Wp + bj + x]);
    return (float)s;
}

/* OptimalDisparity (:40-85) */
float orc_sad_optimal_disparity(const float* sad, int D) {
    float minv = 65535.f, best = 65535.f, sec = sad[0];
    for (int i = 1; i < D; ++i)
        if (minv > sad[i]) {
            minv = sad[i];
            best = (float)i;
        }
    for (int i = 0; i < D; ++i)
        if (minv != sad[i]) sec = sad[i] < sec ? sad[i] : sec;
    if (sec - minv <= 0.01) return 0; /* float difference compared with the double 0.01 (:66) */
    if (best == 0 || best == D - 1) return 0;
    return best; /* the sub-pixel value (:76-81) is computed and discarded */
}

void orc_sad_costs_left(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, float* costs) {
    const int w = winsize + 1, k = 2 * w + 1, W = Wp - 2 * w;
    for (int i = w; i < Hp - w; ++i)
        for (int j = w; j < Wp - w; ++j) {
            float* s = costs + ((size_t)(i - w) * W + (j - w)) * D;
            for (int d = 0; d < D; ++d) /* :121-132 */
                s[d] = (j - w - d < 0) ? s[d - 1] : sad_window(Lp, Rp, Wp, i - w, j - w, i - w, j - w - d, k);
        }
}

void orc_sad_left(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, int32_t* disp) {
    const int w = winsize + 1, k = 2 * w + 1, W = Wp - 2 * w;
    float* s = (float*)malloc((size_t)D * sizeof(float));
    for (int i = w; i < Hp - winsize - 1; ++i) /* :110-116 */
        for (int j = w; j < Wp - winsize - 1; ++j) {
            for (int d = 0; d < D; ++d)
                s[d] = (j - w - d < 0) ? s[d - 1] : sad_window(Lp, Rp, Wp, i - w, j - w, i - w, j - w - d, k);
            disp[(size_t)(i - w) * W + (j - w)] = (int32_t)orc_sad_optimal_disparity(s, D);
        }
    free(s);
}

/* GetPointDepthRight (:141-182): the loops stop one row and one column early (:154-160); GetMinSadIndex (:22-38) */
void orc_sad_right(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, int32_t* disp) {
    const int w = winsize + 1, k = 2 * w + 1, W = Wp - 2 * w;
    float* s = (float*)malloc((size_t)D * sizeof(float));
    for (int i = w; i < Hp - w - 1; ++i)
        for (int j = w; j < Wp - w - 1; ++j) {
            for (int d = 0; d < D; ++d)
                s[d] = (j + d + w + 1 > Wp) ? s[d - 1] : sad_window(Lp, Rp, Wp, i - w, j - w + d, i - w, j - w, k);
            float m = s[0];
            int idx = 0;
            for (int d = 1; d < D; ++d)
                if (s[d] < m) {
                    m = s[d];
                    idx = d;
                }
            disp[(size_t)(i - w) * W + (j - w)] = idx;
        }
    free(s);
}

/* ------------------------------------------------------------------------------------------------ NCC (NCC/NCC.h) */
/* ComputeCost (:15-49): means first, then centred sums, all in double, in the reference's loop order */
double orc_ncc_cost(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int i, int j, int d) {
    (void)H;
    const int k = 2 * winSize + 1, n = k * k;
    const uint8_t* a = L + (size_t)(i - winSize) * W + (j - winSize);
    const uint8_t* b = R + (size_t)(i - winSize) * W + (j - winSize - d);
    double lm = 0, rm = 0, ls = 0, rs = 0, num = 0;
    for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x) { /* the reference indexes (j,i) here; the window is square (:29-30) */
            lm += a[(size_t)x * W + y];
            rm += b[(size_t)x * W + y];
        }
    lm /= n;
    rm /= n;
    for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x) {
            const double dl = a[(size_t)y * W + x] - lm, dr = b[(size_t)y * W + x] - rm;
            ls += pow(dl, 2);
            rs += pow(dr, 2);
            num += dl * dr;
        }
    return num / (sqrt(ls) * sqrt(rs));
}

/* NCC_algorithem (:69-95) with WinTakeAll (:53-67): arg-max whose running maximum is rounded to float; a disparity
 * whose window leaves the image scores `invalid` = 255 (:7,:88), which beats every correlation */
void orc_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int D, int32_t* disp) {
    double* cost = (double*)malloc((size_t)D * sizeof(double));
    for (int i = winSize; i < H - winSize; ++i)
        for (int j = winSize; j < W - winSize; ++j) {
            for (int d = 0; d < D; ++d) cost[d] = (j - winSize - d >= 0) ? orc_ncc_cost(L, R, H, W, winSize, i, j, d) : 255.0;
            int best = 0;
            float m = (float)cost[0];
            for (int d = 1; d < D; ++d)
                if (m < cost[d]) {
                    best = d;
                    m = (float)cost[d];
                }
            disp[(size_t)i * W + j] = best;
        }
    free(cost);
}

/* cost rows of output rows [i0, i1): NCC_algorithem's inner loop (:73-90); double [i1-i0][W][D] */
void orc_ncc_costs(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int D, int i0, int i1, double* out) {
    for (int i = i0; i < i1; ++i)
        for (int j = 0; j < W; ++j) {
            double* row = out + ((size_t)(i - i0) * W + j) * D;
            const int interior = i >= winSize && i < H - winSize && j >= winSize && j < W - winSize;
            for (int d = 0; d < D; ++d)
                row[d] = !interior ? 0.0 : (j - winSize - d >= 0) ? orc_ncc_cost(L, R, H, W, winSize, i, j, d) : 255.0;
        }
}

/* ------------------------------------------------------------------------------------------------ ASW (ASW/ASW.h) */
void orc_asw_masks(int winSize, double spaceSigma, double colorSigma, double* space_out, double* color_out) {
    const int k = 2 * winSize + 3, c = (k - 1) / 2;
    for (int i = 0; i < k; ++i) /* getGausssianMask (:16-35) */
        for (int j = 0; j < k; ++j)
            space_out[i * k + j] = exp(-(pow(j - c, 2) + pow(i - c, 2)) / (2 * spaceSigma * spaceSigma));
    for (int i = 0; i < 256; ++i) color_out[i] = exp(-(i * i) / (2 * colorSigma * colorSigma)); /* getColorMask (:41-47) */
}

/* bilateralfiterWight (:210-257) on two k x k windows */
static float asw_cost(const uint8_t* A, const uint8_t* B, int Wp, int k, int T, const double* space, const double* color) {
    const int c = (k - 1) / 2;
    const int ca = A[(size_t)c * Wp + c], cb = B[(size_t)c * Wp + c];
    double sw = 0, sv = 0;
    for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x) {
            const int pa = A[(size_t)y * Wp + x], pb = B[(size_t)y * Wp + x];
            const double m0 = color[abs(pa - ca)] * space[y * k + x];
            const double m1 = color[abs(pb - cb)] * space[y * k + x];
            const double m2 = m0 * m1;
            int e = abs(pa - pb);
            if (e > T) e = T; /* :359-366 */
            sw += m2;
            sv += m2 * (double)e;
        }
    return (float)(sv / sw);
}

/* AdaptiveSupportWeight (:329-378) / AdaptiveSupportWeightRight (:382-431) + WinTakeAll (:193-208).  In the right view
 * the last winSize+2 output columns are invalid at d = 0 and the reference reads costVolume[-1] there (UB); this
 * restatement leaves those pixels at 0 and parity excludes them. */
void orc_asw(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int winSize, int D, double spaceSigma, double colorSigma,
             int T, int right_view, float* disp) {
    const int wins = winSize + 1, k = 2 * wins + 1, W = Wp - 2 * wins;
    double* space = (double*)malloc((size_t)k * k * sizeof(double));
    double color[256];
    orc_asw_masks(winSize, spaceSigma, colorSigma, space, color);
    float* cost = (float*)malloc((size_t)D * sizeof(float));
    for (int i = wins; i < Hp - wins; ++i)
        for (int j = wins; j < Wp - wins; ++j) {
            int ok = 1;
            for (int d = 0; d < D; ++d) {
                int valid;
                const uint8_t *a, *b;
                if (!right_view) {
                    valid = j - wins - d >= 0;
                    a = Lp + (size_t)(i - wins) * Wp + (j - wins);
                    b = Rp + (size_t)(i - wins) * Wp + (j - wins - d);
                } else {
                    valid = j + wins + d + 1 < Wp - wins;
                    a = Rp + (size_t)(i - wins) * Wp + (j - wins);
                    b = Lp + (size_t)(i - wins) * Wp + (j - wins + d);
                }
                if (valid) cost[d] = asw_cost(a, b, Wp, k, T, space, color);
                else if (d == 0) { ok = 0; break; }
                else cost[d] = cost[d - 1];
            }
            float best = 0;
            if (ok) {
                float m = cost[0];
                for (int d = 1; d < D; ++d)
                    if (m > cost[d]) {
                        best = (float)d;
                        m = cost[d];
                    }
            }
            disp[(size_t)(i - wins) * W + (j - wins)] = best;
        }
    free(space);
    free(cost);
}

/* costVolume rows of the left view for output rows [i0, i1) (:346-372); float [i1-i0][W][D] */
void orc_asw_costs(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int winSize, int D, double spaceSigma, double colorSigma,
                   int T, int i0, int i1, float* out) {
    (void)Hp;
    const int wins = winSize + 1, k = 2 * wins + 1, W = Wp - 2 * wins;
    double* space = (double*)malloc((size_t)k * k * sizeof(double));
    double color[256];
    orc_asw_masks(winSize, spaceSigma, colorSigma, space, color);
    for (int io = i0; io < i1; ++io)
        for (int jo = 0; jo < W; ++jo) {
            const int i = io + wins, j = jo + wins;
            float* row = out + ((size_t)(io - i0) * W + jo) * D;
            const uint8_t* a = Lp + (size_t)(i - wins) * Wp + (j - wins);
            for (int d = 0; d < D; ++d)
                row[d] = (j - wins - d >= 0) ? asw_cost(a, Rp + (size_t)(i - wins) * Wp + (j - wins - d), Wp, k, T, space, color)
                                             : row[d - 1];
        }
    free(space);
}

/* CrossCheckDiaparity (:108-145) */
void orc_asw_crosscheck(const float* dispL, const float* dispR, int H, int W, uint8_t* out) {
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const int lv = (int)dispL[(size_t)i * W + j];
            const float rv = dispR[(size_t)i * W + j - lv];
            const float diff = fabsf((float)lv - rv);
            out[(size_t)i * W + j] = diff > 5 ? 0 : (uint8_t)lv;
        }
}

/* TransformToShow (PostProcessing.h:137-154): `const int v = disp[..]` then (uchar)v; 0 stays 0.  A float outside the
 * int range (+inf, 2147483648.f) is converted with cvttss2si on x86-64, which yields INT_MIN, whose low byte is 0. */
void orc_transform_to_show(const float* disp, int H, int W, uint8_t* out) {
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        const float f = disp[p];
        int v = 0;
        if (f > -2147483648.f && f < 2147483648.f) v = (int)f;
        out[p] = (uint8_t)v;
    }
}

/* FillTheHole (PostProcessing.h:156-248) as written: width/height are swapped (:158-159), "invalid" is 0xffff (:184,:215)
 * although the LR check writes +inf, the 8-direction table switches for good at the first target pixel whose y equals
 * height/2 (:198-200), targets without any hit are filled with 0 (fill_disps is value-initialised, :177,:239-244), and
 * the third phase replaces the mismatch list by all 0xffff pixels (:179-189).  Never called by the reference's main. */
static int cmp_float2(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}
static void fill_the_hole_impl(int width, int height, float invalid, int dispRange, float* disp, const int32_t* occ, int32_t n_occ,
                               int32_t* mis, int32_t* n_mis, int32_t mis_capacity) {
    const float pi = 3.1415926f;
    const float angle1[8] = {pi, 3 * pi / 4, pi / 2, pi / 4, 0, 7 * pi / 4, 3 * pi / 2, 5 * pi / 4};
    const float angle2[8] = {pi, 5 * pi / 4, 3 * pi / 2, 7 * pi / 4, 0, pi / 4, pi / 2, 3 * pi / 4};
    const float* angle = angle1;
    const int max_len = (int)(1.0 * dispRange);
    const size_t cap = (size_t)width * height > (size_t)mis_capacity ? (size_t)width * height : (size_t)mis_capacity;
    int32_t* cur_mis = (int32_t*)malloc((cap + 1) * 2 * sizeof(int32_t));
    int32_t ncur_mis = *n_mis;
    memcpy(cur_mis, mis, (size_t)ncur_mis * 2 * sizeof(int32_t));
    for (int k = 0; k < 3; ++k) {
        const int32_t* trg = k == 0 ? occ : cur_mis;
        int32_t ntrg = k == 0 ? n_occ : ncur_mis;
        if (ntrg == 0) continue;
        float* fill = (float*)calloc((size_t)((k == 2 ? (size_t)width * height : (size_t)ntrg) + 1), sizeof(float));
        if (k == 2) {
            ncur_mis = 0;
            for (int i = 0; i < height; ++i)
                for (int j = 0; j < width; ++j)
                    if (disp[(size_t)i * width + j] == invalid) {
                        cur_mis[2 * (size_t)ncur_mis] = i;
                        cur_mis[2 * (size_t)ncur_mis + 1] = j;
                        ++ncur_mis;
                    }
            trg = cur_mis;
            ntrg = ncur_mis;
        }
        for (int32_t n = 0; n < ntrg; ++n) {
            const int y = trg[2 * n], x = trg[2 * n + 1];
            if (y == height / 2) angle = angle2;
            float coll[8];
            int nc = 0;
            for (int s = 0; s < 8; ++s) {
                const float sina = (float)sin(angle[s]), cosa = (float)cos(angle[s]);
                for (int m = 1; m < max_len; ++m) {
                    const int yy = (int)lround(y + m * sina), xx = (int)lround(x + m * cosa);
                    if (yy < 0 || yy >= height || xx < 0 || xx >= width) break;
                    const float d = disp[(size_t)yy * width + xx];
                    if (d != invalid) {
                        coll[nc++] = d;
                        break;
                    }
                }
            }
            if (nc == 0) continue;
            qsort(coll, (size_t)nc, sizeof(float), cmp_float2);
            fill[n] = k == 0 ? (nc > 1 ? coll[1] : coll[0]) : coll[nc / 2];
        }
        for (int32_t n = 0; n < ntrg; ++n) disp[(size_t)trg[2 * n] * width + trg[2 * n + 1]] = fill[n];
        free(fill);
    }
    *n_mis = ncur_mis;
    for (int32_t k = 0; k < ncur_mis && k < mis_capacity; ++k) {
        mis[2 * k] = cur_mis[2 * k];
        mis[2 * k + 1] = cur_mis[2 * k + 1];
    }
    free(cur_mis);
}
void orc_fill_the_hole(int row, int col, int dispRange, float* disp, const int32_t* occ, int32_t n_occ, int32_t* mis, int32_t* n_mis,
                       int32_t mis_capacity) {
    fill_the_hole_impl(row, col, 65535.f, dispRange, disp, occ, n_occ, mis, n_mis, mis_capacity);  /* :158-159 swap */
}
/* The evident intent (NOT in the reference, parity unpinned): width = col, height = row, invalid = +inf. */
void orc_fill_the_hole_fixed(int row, int col, int dispRange, float* disp, const int32_t* occ, int32_t n_occ, int32_t* mis,
                             int32_t* n_mis, int32_t mis_capacity) {
    fill_the_hole_impl(col, row, INFINITY, dispRange, disp, occ, n_occ, mis, n_mis, mis_capacity);
}

/* ================================================================================================ multi-step refinement
 * PARITY UNPINNED -- the reference has NO implementation of these steps: only the option fields irv_ts, irv_th and
 * do_discontinuity_adjustment exist (CBLSM/adcensus_types.h:58-59,65) and nothing reads them.  BASELINE.json's north_star
 * names "iterative region voting, interpolation and border refinement", so a FIXED SPECIFICATION after Mei et al.,
 * "On building an accurate stereo matching system on graphics hardware" (2011), sections 2.4, is stated here and the CUDA
 * path is checked against this restatement only.  Conventions: disparity maps are float [H][W] with +inf = invalid (what
 * LeftRightConsistency writes, PostProcessing.h:93,131); lists are (row, col) pairs in raster order, as LeftRightConsistency
 * emits them; arms are CrossAggregator's {left,right,top,bottom} bytes (cross_aggregator.h:17-20).  Every phase reads the map
 * as it was when the phase (or the voting iteration) began and writes afterwards, so results do not depend on a visiting
 * order. */

/* Iterative region voting: a listed (invalid) pixel p collects the histogram of the valid disparities in its cross
 * support region (for each row y+t of p's vertical arm: the horizontal arm span of pixel (y+t, x)), count = their number,
 * best = the first most frequent disparity; if count > ts and hist[best] / count > th the pixel takes `best` and leaves
 * its list.  `iters` synchronous iterations.  cls: 0 valid, 1 occlusion, 2 mismatch (raster map of the two lists). */
void orc_refine_region_voting(float* disp, uint8_t* cls, const uint8_t* arms, int H, int W, int D, int ts, float th, int iters) {
    float* next = (float*)malloc((size_t)H * W * sizeof(float));
    int* hist = (int*)malloc((size_t)D * sizeof(int));
    for (int it = 0; it < iters; ++it) {
        memcpy(next, disp, (size_t)H * W * sizeof(float));
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                if (!cls[(size_t)y * W + x]) continue;
                memset(hist, 0, (size_t)D * sizeof(int));
                int count = 0;
                const uint8_t* a = arms + ((size_t)y * W + x) * 4;
                for (int t = -(int)a[2]; t <= (int)a[3]; ++t) {
                    const int yy = y + t;
                    if (yy < 0 || yy >= H) continue;
                    const uint8_t* b = arms + ((size_t)yy * W + x) * 4;
                    for (int s = -(int)b[0]; s <= (int)b[1]; ++s) {
                        const int xx = x + s;
                        if (xx < 0 || xx >= W) continue;
                        const float d = disp[(size_t)yy * W + xx];
                        if (d == INFINITY || !(d >= 0.f)) continue;
                        const int di = (int)lroundf(d);
                        if (di < 0 || di >= D) continue;
                        hist[di]++;
                        count++;
                    }
                }
                if (count <= ts) continue;
                int best = 0;
                for (int d = 1; d < D; ++d)
                    if (hist[d] > hist[best]) best = d;
                if ((float)hist[best] / (float)count > th) next[(size_t)y * W + x] = (float)best;
            }
        for (size_t k = 0; k < (size_t)H * W; ++k)
            if (cls[k] && next[k] != INFINITY) {
                disp[k] = next[k];
                cls[k] = 0;
            }
    }
    free(next);
    free(hist);
}

/* Proper interpolation: a listed pixel looks along 16 directions (angle k*pi/8, tables sina/cosa evaluated by the caller in
 * float) for the first valid disparity, at most max_len - 1 steps, stopping at the image border; an OCCLUSION takes the
 * smallest disparity found (background), a MISMATCH the disparity of the hit whose BGR colour is closest to its own
 * (max-channel distance, first direction among equals).  Pixels without any hit stay invalid. */
void orc_refine_interpolate(float* disp, uint8_t* cls, const uint8_t* bgr, int H, int W, int max_len, const float* sina, const float* cosa) {
    float* next = (float*)malloc((size_t)H * W * sizeof(float));
    memcpy(next, disp, (size_t)H * W * sizeof(float));
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint8_t c = cls[(size_t)y * W + x];
            if (!c) continue;
            float best_d = INFINITY;
            int best_dist = 1 << 30;
            const uint8_t* p = bgr + ((size_t)y * W + x) * 3;
            for (int k = 0; k < 16; ++k)
                for (int m = 1; m < max_len; ++m) {
                    const int yy = (int)lroundf((float)y + (float)m * sina[k]);
                    const int xx = (int)lroundf((float)x + (float)m * cosa[k]);
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) break;
                    const float d = disp[(size_t)yy * W + xx];
                    if (d == INFINITY) continue;
                    if (c == 1) {
                        if (d < best_d) best_d = d;
                    } else {
                        const uint8_t* q = bgr + ((size_t)yy * W + xx) * 3;
                        int dist = abs((int)p[0] - (int)q[0]);
                        const int d1 = abs((int)p[1] - (int)q[1]), d2 = abs((int)p[2] - (int)q[2]);
                        if (d1 > dist) dist = d1;
                        if (d2 > dist) dist = d2;
                        if (dist < best_dist) {
                            best_dist = dist;
                            best_d = d;
                        }
                    }
                    break;
                }
            next[(size_t)y * W + x] = best_d;
        }
    for (size_t k = 0; k < (size_t)H * W; ++k)
        if (cls[k] && next[k] != INFINITY) {
            disp[k] = next[k];
            cls[k] = 0;
        }
    free(next);
}

/* Depth-discontinuity adjustment: on an edge pixel of the disparity map (3x3 Sobel, |gx| + |gy| > thres, interior pixels)
 * with a valid disparity, the disparities of the left and the right neighbour replace it when the aggregated cost volume
 * rates them lower at this pixel (left first, then right against the better of the two). */
void orc_refine_discontinuity(float* disp, const float* cost, int H, int W, int D, float thres) {
    float* next = (float*)malloc((size_t)H * W * sizeof(float));
    memcpy(next, disp, (size_t)H * W * sizeof(float));
    for (int y = 1; y < H - 1; ++y)
        for (int x = 1; x < W - 1; ++x) {
            const float* r0 = disp + (size_t)(y - 1) * W + x;
            const float* r1 = disp + (size_t)y * W + x;
            const float* r2 = disp + (size_t)(y + 1) * W + x;
            const float gx = ((r0[1] - r0[-1]) + (2.f * r1[1] - 2.f * r1[-1])) + (r2[1] - r2[-1]);
            const float gy = ((r2[-1] - r0[-1]) + (2.f * r2[0] - 2.f * r0[0])) + (r2[1] - r0[1]);
            if (!(fabsf(gx) + fabsf(gy) > thres)) continue;
            const float d = r1[0];
            if (d == INFINITY || !(d >= 0.f) || lroundf(d) >= D) continue;
            const float* c = cost + ((size_t)y * W + x) * D;
            float cbest = c[lroundf(d)], dbest = d;
            for (int k = 0; k < 2; ++k) {
                const float d2 = k == 0 ? r1[-1] : r1[1];
                if (d2 == INFINITY || !(d2 >= 0.f) || lroundf(d2) >= D) continue;
                const float c2 = c[lroundf(d2)];
                if (c2 < cbest) {
                    cbest = c2;
                    dbest = d2;
                }
            }
            next[(size_t)y * W + x] = dbest;
        }
    memcpy(disp, next, (size_t)H * W * sizeof(float));
    free(next);
}

/* ================================================================================================ remaining variants */
/* costAggregationVertical (CBLSM/CBLSM.h:1227-1272): the same rectangle mean as V5, column-major accumulation */
void orc_cblsm_aggregate_vertical(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR,
                                  const int32_t* armUp, const int32_t* armDown, int D, int H, int W) {
    const size_t npix = (size_t)H * W;
    float* layer = (float*)malloc(npix * sizeof(float));
    for (int d = 0; d < D; ++d) {
        for (size_t p = 0; p < npix; ++p) layer[p] = vol_in[p * D + d];
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const size_t p = (size_t)i * W + j;
                float value = 0;
                int number = 0;
                for (int a = -armL[p]; a <= armR[p]; ++a)
                    for (int b = -armUp[p]; b <= armDown[p]; ++b) {
                        value = value + layer[(long)(i + b) * W + j + a];
                        ++number;
                    }
                vol_out[p * D + d] = value / number;
            }
    }
    free(layer);
}

/* costAggregationV4 (CBLSM/CBLSM.h:1128-1176): arm arrays per (pixel, d) cell, exclusive upper bounds, row-major;
 * an empty window divides 0 by 0 */
void orc_cblsm_aggregate_v4(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR, const int32_t* armUp,
                            const int32_t* armDown, int D, int H, int W) {
    const size_t npix = (size_t)H * W;
    float* layer = (float*)malloc(npix * sizeof(float));
    for (int d = 0; d < D; ++d) {
        for (size_t p = 0; p < npix; ++p) layer[p] = vol_in[p * D + d];
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const size_t c = ((size_t)i * W + j) * D + d;
                float value = 0;
                int number = 0;
                for (int b = -armUp[c]; b < armDown[c]; ++b)
                    for (int a = -armL[c]; a < armR[c]; ++a) {
                        value = value + layer[(long)(i + b) * W + j + a];
                        ++number;
                    }
                vol_out[c] = value / number;
            }
    }
    free(layer);
}

/* CrossCheckDiaparity on CV_32S maps (SAD/Sad.h:184-221); right[j - lv] is the flat element i*W + j - lv of a contiguous
 * Mat (read as 0 where that lies before the buffer, which the reference leaves undefined) */
void orc_sad_crosscheck(const int32_t* dispL, const int32_t* dispR, int H, int W, int32_t* out, int32_t* occ, int32_t* n_occ,
                        int32_t* mis, int32_t* n_mis) {
    int no = 0, nm = 0;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const long p = (long)i * W + j;
            const int lv = dispL[p];
            const long q = p - lv;
            const int rv = (q >= 0 && q < (long)H * W) ? dispR[q] : 0;
            long diff = (long)lv - (long)rv;
            if (diff < 0) diff = -diff;
            if (diff > 5) {
                if (lv < rv) {
                    occ[2 * no] = i;
                    occ[2 * no + 1] = j;
                    ++no;
                } else {
                    mis[2 * nm] = i;
                    mis[2 * nm + 1] = j;
                    ++nm;
                }
                out[p] = 2147483647; /* (int)Invalid_Float under g++ */
            } else {
                out[p] = lv;
            }
        }
    *n_occ = no;
    *n_mis = nm;
}

/* the sub-pixel value OptimalDisparity computes and discards (SAD/Sad.h:76-84); 0 where the function returns 0.
 * NOT observable in the reference (it returns best_disp): a restatement of those lines only. */
float orc_sad_subpixel(const float* sad, int n) {
    float minv = 65535.f, best = 65535.f, sec = sad[0];
    for (int i = 1; i < n; ++i)
        if (minv > sad[i]) {
            minv = sad[i];
            best = (float)i;
        }
    for (int i = 0; i < n; ++i)
        if (minv != sad[i] && sad[i] < sec) sec = sad[i];
    if (sec - minv <= 0.01) return 0.f;
    if (best == 0 || best == n - 1) return 0.f;
    const float c1 = sad[(int)best - 1], c2 = sad[(int)best + 1];
    float divided = c1 + c2 - 2 * minv;
    if (divided < 1.f) divided = 1.f;
    return best + (c1 - c2) / (2 * divided);
}

/* ncc() (NCC/NCC.h:117-272): shifted image, clipped 11x11 window, the function's own `n`, doubles in its loop order */
void orc_ncc_shift(const uint8_t* L, const uint8_t* R0, int H, int W, int type_right, int add_constant, uint8_t* depth) {
    const int max_offset = 79, ks = 5;
    uint8_t* R = (uint8_t*)malloc((size_t)H * W);
    uint8_t* tmp = (uint8_t*)malloc((size_t)H * W);
    double* best = (double*)malloc((size_t)H * W * sizeof(double));
    for (size_t k = 0; k < (size_t)H * W; ++k) {
        int v = R0[k] + (add_constant ? 10 : 0);
        R[k] = (uint8_t)(v > 255 ? 255 : v);
        best[k] = -2;
        depth[k] = 0;
    }
    for (int off = 1; off <= max_offset; ++off) {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                if (!type_right) tmp[(size_t)y * W + x] = x < off ? R[(size_t)y * W + x] : R[(size_t)y * W + x - off];
                else tmp[(size_t)y * W + x] = x < W - off ? L[(size_t)y * W + x + off] : L[(size_t)y * W + x];
            }
        const uint8_t* A = type_right ? tmp : L;
        const uint8_t* B = type_right ? R : tmp;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int sx = x - ks > 0 ? x - ks : 0, sy = y - ks > 0 ? y - ks : 0;
                const int ex = x + ks < W - 1 ? x + ks : W - 1, ey = y + ks < H - 1 ? y + ks : H - 1;
                const double n = (double)((ey - sy) * (ex - sx));
                double ma = 0, mb = 0, sa = 0, sb = 0, num = 0;
                for (int i = sy; i <= ey; ++i)
                    for (int j = sx; j <= ex; ++j) {
                        ma += A[(size_t)i * W + j];
                        mb += B[(size_t)i * W + j];
                    }
                ma /= n;
                mb /= n;
                for (int i = sy; i <= ey; ++i)
                    for (int j = sx; j <= ex; ++j) {
                        sa += pow(A[(size_t)i * W + j] - ma, 2);
                        sb += pow(B[(size_t)i * W + j] - mb, 2);
                        num += (A[(size_t)i * W + j] - ma) * (B[(size_t)i * W + j] - mb);
                    }
                num /= n;
                sa /= n;
                sb /= n;
                const double r = num / (sqrt(sa) * sqrt(sb)) / n;
                if (r > best[(size_t)y * W + x]) {
                    best[(size_t)y * W + x] = r;
                    depth[(size_t)y * W + x] = (uint8_t)(off * 3);
                }
            }
    }
    free(R);
    free(tmp);
    free(best);
}

/* ------------------------------------------------------------------------------------------------------------------
 * The ASW driver's display chain (ASW/ASWeight.cpp:66-78): normalize(lastDisp, 0, 255, NORM_MINMAX) on the 8-bit
 * cross-checked map, filterSpeckles(lastDisp, 0, 40, 2), medianBlur 5, FillImageNew, medianBlur 3.  Three of these are
 * OpenCV functions: the sources are not in the reference tree (the project links a prebuilt OpenCV), so they are restated
 * from OpenCV's published algorithms and pinned by golden vectors made with the cv2 4.13 wheel of the build container
 * (oracle/make_golden_display.py); FillImageNew is pinned against the compiled reference. */

/* cv::normalize(src, dst, 0, 255, NORM_MINMAX) for CV_8U -> CV_8U: scale = 255 * (1 / (max - min)) in double (0 for a
 * constant image), shift = -min * scale; convertTo evaluates src * (float)scale + (float)shift in float with ONE rounding
 * (the SIMD path multiplies and adds fused) and saturate_cast<uchar> rounds half to even. */
void orc_normalize_minmax_u8_from_u8(const uint8_t* src, int H, int W, uint8_t* out) {
    const size_t n = (size_t)H * W;
    int lo = 255, hi = 0;
    for (size_t k = 0; k < n; ++k) {
        if (src[k] < lo) lo = src[k];
        if (src[k] > hi) hi = src[k];
    }
    const double scale = 255.0 * ((double)(hi - lo) > 2.220446049250313e-16 ? 1.0 / (double)(hi - lo) : 0.0);
    const double shift = 0.0 - (double)lo * scale;
    const float a = (float)scale, b = (float)shift;
    for (size_t k = 0; k < n; ++k) {
        const float r = fmaf((float)src[k], a, b);
        long q = lrintf(r);
        out[k] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
}

/* cv::filterSpeckles(img, newVal, maxSpeckleSize, maxDiff) on CV_8UC1 (calib3d): the connected components of the graph
 * {4-neighbours, both != newVal, |a - b| <= maxDiff}; a component of at most maxSpeckleSize pixels is set to newVal. */
void orc_filter_speckles_u8(uint8_t* img, int H, int W, int new_val, int max_size, int max_diff) {
    const size_t n = (size_t)H * W;
    int32_t* label = (int32_t*)calloc(n, sizeof(int32_t));
    int32_t* stack = (int32_t*)malloc(n * sizeof(int32_t));
    uint8_t* small = (uint8_t*)malloc(n + 1);
    int32_t cur = 0;
    for (size_t p0 = 0; p0 < n; ++p0) {
        if (img[p0] == new_val) continue;
        if (label[p0]) {
            if (small[label[p0]]) img[p0] = (uint8_t)new_val;
            continue;
        }
        int32_t top = 0, count = 0;
        stack[top++] = (int32_t)p0;
        label[p0] = ++cur;
        while (top) {
            const int32_t p = stack[--top];
            const int i = p / W, j = p - i * W;
            const int v = img[p];
            ++count;
            const int32_t nb[4] = {j + 1 < W ? p + 1 : -1, j > 0 ? p - 1 : -1, i + 1 < H ? p + W : -1, i > 0 ? p - W : -1};
            for (int t = 0; t < 4; ++t) {
                const int32_t q = nb[t];
                if (q < 0 || label[q] || img[q] == new_val || abs(v - (int)img[q]) > max_diff) continue;
                label[q] = cur;
                stack[top++] = q;
            }
        }
        /* pixels of a small component that were already set to newVal are skipped above by the first test; the values
         * read by the flood fill of LATER components are therefore the modified ones, as in OpenCV -- it makes no
         * difference, a pixel equal to newVal is never an edge endpoint either way */
        small[cur] = count <= max_size;
        if (small[cur]) img[p0] = (uint8_t)new_val;
    }
    free(label);
    free(stack);
    free(small);
}

/* cv::medianBlur on CV_8UC1, ksize 3 or 5: the median of the k*k window with BORDER_REPLICATE. */
void orc_median_blur_u8(const uint8_t* in, uint8_t* out, int H, int W, int k) {
    const int r = k / 2;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            int hist[256] = {0};
            for (int di = -r; di <= r; ++di)
                for (int dj = -r; dj <= r; ++dj) {
                    const int y = i + di < 0 ? 0 : i + di >= H ? H - 1 : i + di;
                    const int x = j + dj < 0 ? 0 : j + dj >= W ? W - 1 : j + dj;
                    ++hist[in[(size_t)y * W + x]];
                }
            int acc = 0, v = 0;
            while ((acc += hist[v]) <= k * k / 2) ++v;
            out[(size_t)i * W + j] = (uint8_t)v;
        }
}

/* FillImageNew (ASW/ASW.h:434-511) as written.  Every zero pixel, in raster order and reading the UNMODIFIED image:
 *   to the left it probes columns c, c-1, c-3, c-6, ... (the offset grows AND accumulates, :468, :483) and takes the first
 *   non-zero value; if that runs off the row, to the right c, c+1, c+3, ... where column == cols is still read (:489 tests
 *   `> col`): that is the first pixel of the next row, and `past_end` for the last row.  Running beyond pushes 0; ending
 *   exactly on column == cols with a zero there pushes NOTHING (:485 leaves the loop), so from then on pixel i receives
 *   the value found for a later pixel (:503-509 pairs the two vectors by index) and the last pixels read beyond the value
 *   vector -- those are left untouched here and reported through n_unassigned. */
void orc_fill_image_new(uint8_t* disp, int H, int W, int past_end, int32_t* n_unassigned) {
    const size_t n = (size_t)H * W;
    int32_t* vals = (int32_t*)malloc((n + 1) * sizeof(int32_t));
    size_t nv = 0, nz = 0;
    for (size_t p = 0; p < n; ++p) {
        if (disp[p]) continue;
        ++nz;
        const int i = (int)(p / W), c = (int)(p - (size_t)i * W);
        const uint8_t* row = disp + (size_t)i * W;
        int found = -1;
        for (int col = c, off = 0; col >= 0; ++off, col -= off)
            if (row[col]) { found = row[col]; break; }
        if (found < 0) {
            int col = c, off = 0, pushed = 0;
            while (col < W) {
                col += off;
                if (col > W) { found = 0; pushed = 1; break; }
                const int v = (col == W && i == H - 1) ? past_end : row[col];
                if (v) { found = v; pushed = 1; break; }
                ++off;
            }
            if (!pushed) continue;
        }
        vals[nv++] = found;
    }
    /* the image is written only now, after every search (:503-509); walk a snapshot of the zero mask */
    uint8_t* zero = (uint8_t*)malloc(n);
    for (size_t p = 0; p < n; ++p) zero[p] = disp[p] == 0;
    size_t k = 0;
    for (size_t p = 0; p < n; ++p) {
        if (!zero[p]) continue;
        if (k < nv) disp[p] = (uint8_t)vals[k];
        ++k;
    }
    if (n_unassigned) *n_unassigned = (int32_t)(nz - nv);
    free(zero);
    free(vals);
}
