// Window matchers: SAD (SAD/Sad.h), NCC (NCC/NCC.h) and adaptive support weight (ASW/ASW.h).
//
//   SAD / NCC  the per-disparity window sum (|l-r| resp. l*r) is a separable box sum.  box_sum_kernel stages a band of
//              image rows in shared memory, each thread owns one (column, disparity) pair, forms the horizontal
//              (2w+1)-tap sum of a row from shared memory and slides the vertical window with a ring of the last 2w+1
//              row sums -- O(2w+1) taps per cell instead of (2w+1)^2, all in exact integer arithmetic.  NCC's other four
//              sums (sum l, sum l^2, sum r, sum r^2) do not depend on d and are per-pixel maps.
//   ASW        bilateral support weights are not separable: 625 taps per cell from shared-memory tiles, with the
//              reference pixel's weight table computed once per pixel and reused for all disparities.
#include <cmath>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------- separable box sums
constexpr int BOX_TX = 4, BOX_TD = 64, BOX_RB = 64;  // columns x disparities per block, rows per band

// A: the view's own padded image, B: the other image, both [Hp][Wp] uint8.  sign -1: other window at column -d
// (left view), +1: at column +d (right view).  out: int32 [Ho][Wo][D], Ho = Hp-2w, Wo = Wp-2w; cells whose other
// window leaves the image get -1 (Sad.h:125,:167; NCC.h:81).
template <int MODE>  // 0: |a-b| (Sad.h:15-20), 1: a*b
__global__ void __launch_bounds__(BOX_TX* BOX_TD) box_sum_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                                                  int Hp, int Wp, int w, int sign, int D, int* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = 2 * w + 1;
    const int Ho = Hp - 2 * w, Wo = Wp - 2 * w;
    const int own_w = BOX_TX + 2 * w, oth_w = BOX_TX + 2 * w + BOX_TD - 1, rows = BOX_RB + 2 * w;
    int* ring = reinterpret_cast<int*>(smem_raw);                                   // [K][256]
    uint8_t* s_own = reinterpret_cast<uint8_t*>(ring + (size_t)K * BOX_TX * BOX_TD);  // [rows][own_w]
    uint8_t* s_oth = s_own + (size_t)rows * own_w;                                  // [rows][oth_w]

    const int jo0 = blockIdx.x * BOX_TX, io0 = blockIdx.y * BOX_RB, d0 = blockIdx.z * BOX_TD;
    const int tid = threadIdx.x, xl = tid / BOX_TD, dl = tid % BOX_TD;
    const int oth_base = sign < 0 ? jo0 - (d0 + BOX_TD - 1) : jo0 + d0;  // padded column of s_oth[.][0]
    for (int t = tid; t < rows * own_w; t += blockDim.x) {
        const int r = t / own_w, c = t - r * own_w;
        const int y = io0 + r, x = jo0 + c;
        s_own[t] = (y < Hp && x < Wp) ? A[(size_t)y * Wp + x] : 0;
    }
    for (int t = tid; t < rows * oth_w; t += blockDim.x) {
        const int r = t / oth_w, c = t - r * oth_w;
        const int y = io0 + r, x = oth_base + c;
        s_oth[t] = (y < Hp && x >= 0 && x < Wp) ? B[(size_t)y * Wp + x] : 0;
    }
    __syncthreads();

    const int jo = jo0 + xl, d = d0 + dl;
    const int oth_off = sign < 0 ? xl + (BOX_TD - 1 - dl) : xl + dl;
    const bool valid = sign < 0 ? (jo - d >= 0) : (jo + 2 * w + d <= Wp - 1);
    const bool in_range = jo < Wo && d < D;
    int acc = 0;
    for (int pr = 0; pr < rows; ++pr) {
        const uint8_t* a = s_own + (size_t)pr * own_w + xl;
        const uint8_t* b = s_oth + (size_t)pr * oth_w + oth_off;
        int hs = 0;
        for (int t = 0; t < K; ++t) {
            const int va = a[t], vb = b[t];
            hs += MODE == 0 ? abs(va - vb) : va * vb;
        }
        int* slot = ring + (size_t)(pr % K) * (BOX_TX * BOX_TD) + tid;
        if (pr >= K) acc -= *slot;
        acc += hs;
        *slot = hs;
        const int io = io0 + pr - (K - 1);
        if (pr >= K - 1 && io < Ho && in_range) out[((size_t)io * Wo + jo) * D + d] = valid ? acc : -1;
    }
}

static smt_status_t k_box_sum(smt_ctx* ctx, int mode, const uint8_t* A, const uint8_t* B, int Hp, int Wp, int w, int sign, int D,
                              int* out) {
    const int K = 2 * w + 1, Ho = Hp - 2 * w, Wo = Wp - 2 * w;
    const size_t smem = sizeof(int) * (size_t)K * BOX_TX * BOX_TD + (size_t)(BOX_RB + 2 * w) * (BOX_TX + 2 * w) +
                        (size_t)(BOX_RB + 2 * w) * (BOX_TX + 2 * w + BOX_TD - 1);
    if (smem > 200 * 1024) return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "window radius %d too large for the box-sum tile", w);
    dim3 grid(cdiv(Wo, BOX_TX), cdiv(Ho, BOX_RB), cdiv(D, BOX_TD));
    if (mode == 0) {
        if (smem > 48 * 1024) SMT_CUDA(ctx, cudaFuncSetAttribute(box_sum_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMT_LAUNCH(ctx, box_sum_kernel<0>, grid, BOX_TX * BOX_TD, smem, A, B, Hp, Wp, w, sign, D, out);
    } else {
        if (smem > 48 * 1024) SMT_CUDA(ctx, cudaFuncSetAttribute(box_sum_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMT_LAUNCH(ctx, box_sum_kernel<1>, grid, BOX_TX * BOX_TD, smem, A, B, Hp, Wp, w, sign, D, out);
    }
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- SAD winner selection
// One thread per output pixel.  Left view: OptimalDisparity (Sad.h:40-85); right view: GetMinSadIndex (:22-38) and
// the loops that stop one row / column early (:154-160).  costs (optional): float [Ho][Wo][D] with the reference's
// copy-previous rule for invalid disparities (:125-129).
// subpix (optional, left view): the value OptimalDisparity computes from the parabola through the winner's neighbours and then
// discards (Sad.h:76-84: `return best_disp`, not final_disp); 0 wherever the function returns 0.
__global__ void __launch_bounds__(128) sad_select_kernel(const int* __restrict__ vol, int Ho, int Wo, int D, int right_view,
                                                         int* __restrict__ disp, float* __restrict__ costs, float* __restrict__ subpix) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)Ho * Wo) return;
    const int io = (int)(p / Wo), jo = (int)(p - (size_t)io * Wo);
    const int* s = vol + p * D;
    int result = 0;
    if (!right_view) {
        float minv = 65535.f, best = 65535.f, last = 0.f;
        const float s0 = (float)s[0];
        for (int i = 0; i < D; ++i) {
            const int raw = s[i];
            const float v = raw >= 0 ? (float)raw : last;
            last = v;
            if (costs) costs[p * D + i] = v;
            if (i >= 1 && minv > v) {
                minv = v;
                best = (float)i;
            }
        }
        float sec = s0;
        last = 0.f;
        for (int i = 0; i < D; ++i) {
            const int raw = s[i];
            const float v = raw >= 0 ? (float)raw : last;
            last = v;
            if (minv != v) sec = fminf(sec, v);
        }
        if ((double)(sec - minv) <= 0.01)
            result = 0;
        else if (best == 0.f || best == (float)(D - 1))
            result = 0;
        else
            result = (int)best;
        if (subpix) {
            float fd = 0.f;
            if (result != 0) {
                // sad[] as the reference sees it: invalid disparities repeat the previous value (:125-129)
                float c1 = 0.f, c2 = 0.f, lastv = 0.f;
                for (int i = 0; i <= result + 1; ++i) {
                    const int raw = s[i];
                    const float v = raw >= 0 ? (float)raw : lastv;
                    lastv = v;
                    if (i == result - 1) c1 = v;
                    if (i == result + 1) c2 = v;
                }
                const float divided = fmaxf(1.f, __fsub_rn(__fadd_rn(c1, c2), __fmul_rn(2.f, minv)));   // :81
                fd = __fadd_rn(best, __fdiv_rn(__fsub_rn(c1, c2), __fmul_rn(2.f, divided)));            // :82
            }
            subpix[p] = fd;
        }
    } else {
        if (io >= Ho - 1 || jo >= Wo - 1) {
            disp[p] = 0;  // never written by the reference: stays Mat::zeros
            return;
        }
        float m = (float)s[0];
        int idx = 0;
        for (int i = 1; i < D; ++i) {
            const int raw = s[i];
            if (raw >= 0 && (float)raw < m) {
                m = (float)raw;
                idx = i;
            }
        }
        result = idx;
    }
    disp[p] = result;
}

// ---------------------------------------------------------------------------------------------- NCC
// per-pixel window sums of v and v^2 (window centred on the pixel, only where it fits)
__global__ void __launch_bounds__(128) ncc_stats_kernel(const uint8_t* __restrict__ img, int H, int W, int w, int* __restrict__ s1,
                                                        int* __restrict__ s2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= W) return;
    int a = 0, b = 0;
    if (i >= w && i < H - w && j >= w && j < W - w) {
        for (int y = -w; y <= w; ++y)
            for (int x = -w; x <= w; ++x) {
                const int v = img[(size_t)(i + y) * W + j + x];
                a += v;
                b += v * v;
            }
    }
    s1[(size_t)i * W + j] = a;
    s2[(size_t)i * W + j] = b;
}

// One warp per interior pixel.  cost(d) = (n*Slr - Sl*Sr) / (sqrt(n*Sll - Sl^2) * sqrt(n*Srr - Sr^2)) from exact
// integers, then NCC.h:53-67: best = 0, m = float(cost[0]); for d >= 1: if (double)m < cost[d] { best = d; m = float(cost[d]) }.
// m always equals float(running maximum), so d is accepted iff cost[d] > (double)float(max_{k<d} cost[k]) and the answer
// is the LAST accepted d: a prefix-max scan instead of a sequential fold.  An out-of-image window scores 255 (:7,:88).
__global__ void __launch_bounds__(256) ncc_select_kernel(const int* __restrict__ slr, const int* __restrict__ l1,
                                                         const int* __restrict__ l2, const int* __restrict__ r1,
                                                         const int* __restrict__ r2, int H, int W, int w, int D,
                                                         int* __restrict__ disp, double* __restrict__ costs) {
    const int Ho = H - 2 * w, Wo = W - 2 * w;
    const size_t q = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= (size_t)Ho * Wo) return;
    const int lane = threadIdx.x & 31;
    const int io = (int)(q / Wo), jo = (int)(q - (size_t)io * Wo);
    const int i = io + w, j = jo + w;
    const long long n = (long long)(2 * w + 1) * (2 * w + 1);
    const long long sl = l1[(size_t)i * W + j], sll = l2[(size_t)i * W + j];
    const double sdl = sqrt((double)(n * sll - sl * sl));
    double run = 0.0;  // running maximum over the chunks already processed
    bool have = false, nan0 = false;
    int best = 0;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        double c = -INFINITY;
        bool isnan_c = false;
        if (d < D) {
            const int raw = slr[q * D + d];
            if (raw < 0) {
                c = 255.0;
            } else {
                const long long sr = r1[(size_t)i * W + j - d], srr = r2[(size_t)i * W + j - d];
                const double num = (double)(n * (long long)raw - sl * sr);
                c = num / (sdl * sqrt((double)(n * srr - sr * sr)));
                if (costs) costs[((size_t)i * W + j) * D + d] = c;  // NaN for a zero-variance window, like the reference's 0/0
                if (c != c) {
                    isnan_c = true;
                    c = -INFINITY;  // a NaN never compares greater (NCC.h:59) and never raises the maximum
                }
            }
            if (costs && raw < 0) costs[((size_t)i * W + j) * D + d] = 255.0;
        }
        if (d0 == 0) nan0 = __shfl_sync(0xffffffffu, (int)isnan_c, 0) != 0;
        // exclusive prefix maximum inside the chunk, seeded with the running maximum
        double incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = fmax(incl, up);
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = -INFINITY;
        const bool seeded = have || lane > 0;
        if (have) excl = fmax(excl, run);
        const bool acc = d < D && d >= 1 && seeded && !isnan_c && c > (double)(float)excl;
        const unsigned m = __ballot_sync(0xffffffffu, acc);
        if (m) best = d0 + 31 - __clz(m);
        const double cm = __shfl_sync(0xffffffffu, incl, 31);
        run = have ? fmax(run, cm) : cm;
        have = true;
    }
    if (nan0) best = 0;  // m = NaN from the start: no later disparity is ever accepted
    if (lane == 0) disp[(size_t)i * W + j] = best;
}

// ---------------------------------------------------------------------------------------------- ASW
constexpr int ASW_TX = 4, ASW_TD = 32, ASW_Q = 4;  // columns per block, disparity lanes, disparities per lane

// Lp/Rp: replicate-padded grey [Hp][Wp], pad = wins = winSize+1 (ASWeight.cpp:54-55).  One block per (output row,
// ASW_TX columns); thread = (column, disparity lane), each lane owns ASW_Q consecutive disparities.
// cost = sum(wOwn*wOth*min(|a-b|,T)) / sum(wOwn*wOth) with w = colour[|p - centre|] * space[tap] (ASW.h:210-257), taps
// accumulated in the reference's row-major order per cell; first-minimum WTA (:193-208).
// The four windows of a lane in the other image are one pixel apart: along a tap row they read one sequence of pixels,
// cell slot s at tap c taking element c + s (slot s = disparity d0 + s in the right view, d0 + 3 - s in the left view).
// The sequence is loaded four elements at a time into a register window (one new byte per tap instead of four), and the
// own pixel and the own-window weight of a tap are read once for four cells.  The colour table is replicated per lane
// (index v*32 + lane): the lookups are data-dependent and would otherwise collide in the banks.
// Per cell and tap: |q - centre| -> table address -> LDS, one multiply (the spatial weight of the other window is folded
// into the own-window table: wOwn*space*space), |a - q| and min in float, one add, one fused multiply-add.
// shared-memory load from a 32-bit shared address: lets the table lookup below be ONE multiply-add (index * 128 + lane base)
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
constexpr int ASW_PAD = ASW_Q - 1;  // columns of padding on both sides of a staged row of the other image: no clamps in the tap loop

template <bool RIGHT>
__global__ void __launch_bounds__(ASW_TX* ASW_TD) asw_kernel(const uint8_t* __restrict__ own_img, const uint8_t* __restrict__ oth_img,
                                                              int Hp, int Wp, int wins, int D, int T,
                                                              const float* __restrict__ space, const float* __restrict__ color,
                                                              float* __restrict__ disp, float* __restrict__ costs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = 2 * wins + 1, kk = k * k;
    const int Wo = Wp - 2 * wins;
    const int own_w = ASW_TX + k - 1, oth_w = ASW_TX + k - 1 + D - 1, oth_wp = oth_w + 2 * ASW_PAD + ASW_Q;
    float* s_colr = reinterpret_cast<float*>(smem_raw);        // [256][32]
    float* s_wown = s_colr + 256 * 32;                         // [ASW_TX][kk]: colour * space * space
    float* s_ownf = s_wown + ASW_TX * kk;                      // [k][own_w] own pixels as float
    float* s_bestv = s_ownf + k * own_w;                       // [ASW_TX][ASW_TD]
    int* s_bestd = reinterpret_cast<int*>(s_bestv + ASW_TX * ASW_TD);
    uint8_t* s_oth = reinterpret_cast<uint8_t*>(s_bestd + ASW_TX * ASW_TD);  // [k][oth_wp]

    const int io = blockIdx.y, jo0 = blockIdx.x * ASW_TX;
    const int tid = threadIdx.x, xl = tid / ASW_TD, dl = tid % ASW_TD, l32 = tid & 31;
    const int oth_base = RIGHT ? jo0 : jo0 - (D - 1);  // padded-image column of staged column ASW_PAD
    for (int t = tid; t < 256 * 32; t += blockDim.x) s_colr[t] = color[t >> 5];
    for (int t = tid; t < k * own_w; t += blockDim.x) {
        const int r = t / own_w, c = t - r * own_w;
        const int x = jo0 + c;
        s_ownf[t] = x < Wp ? (float)own_img[(size_t)(io + r) * Wp + x] : 0.f;
    }
    for (int t = tid; t < k * oth_wp; t += blockDim.x) {
        const int r = t / oth_wp, c = t - r * oth_wp;
        const int x = oth_base + c - ASW_PAD;
        s_oth[t] = (x >= 0 && x < Wp) ? oth_img[(size_t)(io + r) * Wp + x] : 0;
    }
    __syncthreads();
    // weight table of each reference pixel: colour[|p - centre|] * space, times the spatial weight of the other window
    for (int t = tid; t < ASW_TX * kk; t += blockDim.x) {
        const int x = t / kk, tap = t - x * kk;
        const int r = tap / k, c = tap - r * k;
        const float centre = s_ownf[(size_t)wins * own_w + x + wins];
        const float sp = space[tap];
        s_wown[t] = __fmul_rn(__fmul_rn(color[(int)fabsf(s_ownf[(size_t)r * own_w + x + c] - centre)], sp), sp);
    }
    __syncthreads();

    const int jo = jo0 + xl;
    const unsigned colr_addr = (unsigned)__cvta_generic_to_shared(s_colr + l32);
    const float Tf = (float)T;
    float bestv = INFINITY;
    int bestd = 0x7fffffff;
    for (int d0 = dl * ASW_Q; d0 < D; d0 += ASW_TD * ASW_Q) {
        if (jo >= Wo) break;
        // staged column of element 0 of the lane's pixel sequence (left edge of slot 0's window)
        const int B0 = ASW_PAD + (RIGHT ? xl + d0 : xl + (D - 1 - d0) - (ASW_Q - 1));
        int cb[ASW_Q];
#pragma unroll
        for (int i = 0; i < ASW_Q; ++i) cb[i] = s_oth[(size_t)wins * oth_wp + B0 + i + wins];
        const float* wo = s_wown + xl * kk;
        float sw[ASW_Q], sv[ASW_Q];
#pragma unroll
        for (int i = 0; i < ASW_Q; ++i) sw[i] = sv[i] = 0.f;
        for (int r = 0; r < k; ++r) {
            const float* a = s_ownf + (size_t)r * own_w + xl;
            const uint8_t* b = s_oth + (size_t)r * oth_wp + B0;
            const float* w0 = wo + r * k;
            // The reference accumulates both sums in double (ASW.h:212-255).  float32 accumulators are kept within the
            // 1e-5 bar by summing each tap row on its own and adding the row totals: <= (k-1) + (k-1) roundings on
            // positive terms instead of k*k - 1, i.e. a worst-case relative error of 48 * 2^-24 = 2.9e-6 for k = 25.
            float rw[ASW_Q], rv[ASW_Q];
#pragma unroll
            for (int i = 0; i < ASW_Q; ++i) rw[i] = rv[i] = 0.f;
            int q[2 * ASW_Q - 1];
            float qf[2 * ASW_Q - 1];
#pragma unroll
            for (int i = 0; i < ASW_Q - 1; ++i) {
                q[i] = b[i];
                qf[i] = (float)q[i];
            }
#define SMT_ASW_TAP(C, S0)                                                                  \
    {                                                                                       \
        const float pa = a[C], w = w0[C];                                                   \
        _Pragma("unroll") for (int i = 0; i < ASW_Q; ++i) {                                 \
            const float m2 = __fmul_rn(lds_f32(__usad((unsigned)q[(S0) + i], (unsigned)cb[i], 0u) * 128u + colr_addr), w); \
            const float e = fminf(fabsf(pa - qf[(S0) + i]), Tf);                            \
            rw[i] = __fadd_rn(rw[i], m2);                                                   \
            rv[i] = __fmaf_rn(m2, e, rv[i]);                                                \
        }                                                                                   \
    }
            int c = 0;
            for (; c + ASW_Q <= k; c += ASW_Q) {
#pragma unroll
                for (int i = 0; i < ASW_Q; ++i) {
                    q[ASW_Q - 1 + i] = b[c + ASW_Q - 1 + i];
                    qf[ASW_Q - 1 + i] = (float)q[ASW_Q - 1 + i];
                }
#pragma unroll
                for (int s4 = 0; s4 < ASW_Q; ++s4) SMT_ASW_TAP(c + s4, s4)
#pragma unroll
                for (int i = 0; i < ASW_Q - 1; ++i) {
                    q[i] = q[ASW_Q + i];
                    qf[i] = qf[ASW_Q + i];
                }
            }
            for (; c < k; ++c) {
                q[ASW_Q - 1] = b[c + ASW_Q - 1];
                qf[ASW_Q - 1] = (float)q[ASW_Q - 1];
                SMT_ASW_TAP(c, 0)
#pragma unroll
                for (int i = 0; i < ASW_Q - 1; ++i) {
                    q[i] = q[i + 1];
                    qf[i] = qf[i + 1];
                }
            }
#undef SMT_ASW_TAP
#pragma unroll
            for (int i = 0; i < ASW_Q; ++i) {
                sw[i] = __fadd_rn(sw[i], rw[i]);
                sv[i] = __fadd_rn(sv[i], rv[i]);
            }
        }
        // ascending d inside the thread: strict '<' keeps the first minimum
#pragma unroll
        for (int dd = 0; dd < ASW_Q; ++dd) {
            const int i = RIGHT ? dd : ASW_Q - 1 - dd;  // slot of disparity d0 + dd
            const int d = d0 + dd;
            const bool valid = d < D && (RIGHT ? (jo + d < Wo - wins - 1) : (jo - d >= 0));
            if (valid) {
                const float cost = __fdiv_rn(sv[i], sw[i]);
                if (costs) costs[((size_t)io * Wo + jo) * D + d] = cost;
                if (cost < bestv) {
                    bestv = cost;
                    bestd = d;
                }
            }
        }
    }

    s_bestv[tid] = bestv;
    s_bestd[tid] = bestd;
    __syncthreads();
    // invalid disparities repeat the previous cost (ASW.h:371): the valid ones are a prefix d < nv of the range (written
    // above by threads of this block, visible after the barrier)
    if (costs && jo < Wo) {
        const int nv = RIGHT ? max(0, min(D, Wo - wins - 1 - jo)) : min(D, jo + 1);
        if (nv > 0) {
            const float lastc = costs[((size_t)io * Wo + jo) * D + nv - 1];
            for (int d = nv + dl; d < D; d += ASW_TD) costs[((size_t)io * Wo + jo) * D + d] = lastc;
        }
    }
    if (dl == 0 && jo < Wo) {
        float bv = INFINITY;
        int bd = 0x7fffffff;
        for (int t = 0; t < ASW_TD; ++t) {
            const float v = s_bestv[xl * ASW_TD + t];
            const int dd = s_bestd[xl * ASW_TD + t];
            if (dd != 0x7fffffff && (v < bv || (v == bv && dd < bd))) {
                bv = v;
                bd = dd;
            }
        }
        disp[(size_t)io * Wo + jo] = bd == 0x7fffffff ? 0.f : (float)bd;
    }
}

// CrossCheckDiaparity (ASW.h:108-145)
__global__ void __launch_bounds__(256) asw_crosscheck_kernel(const float* __restrict__ dl, const float* __restrict__ dr, int H, int W,
                                                             uint8_t* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W) return;
    const int j = (int)(p % W);
    const int lv = (int)dl[p];
    const int jj = min(max(j - lv, 0), W - 1);  // the reference indexes j - lv unchecked; left disparities satisfy lv <= j
    const float rv = dr[p - j + jj];
    out[p] = fabsf((float)lv - rv) > 5.f ? 0 : (uint8_t)lv;
}

smt_status_t k_asw_crosscheck(smt_ctx* ctx, const float* dl, const float* dr, int H, int W, uint8_t* out) {
    StageTimer t(ctx, ST_POST);
    SMT_LAUNCH(ctx, asw_crosscheck_kernel, cdiv((long long)H * W, 256), 256, 0, dl, dr, H, W, out);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- SAD cross-check (int maps)
// CrossCheckDiaparity (SAD/Sad.h:184-221): rv = right[i][j - lv]; |lv - rv| > 5 -> invalid, occlusion if lv < rv else
// mismatch; the marker written into the CV_32S map is (int)Invalid_Float, 2147483647 under g++.
// The reference indexes j - lv unchecked; for a contiguous Mat that is the flat element i*cols + j - lv, which is what
// is read here whenever it lies inside the map (it only leaves it in row 0, where the reference reads before the buffer:
// those pixels take rv = 0).  Lists come out in raster order through the LR check's compaction kernels.
__global__ void __launch_bounds__(256) sad_crosscheck_kernel(const int* __restrict__ dl, const int* __restrict__ dr, int H, int W,
                                                             int* __restrict__ out, unsigned char* __restrict__ cls) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W) return;
    const int lv = dl[p];
    const long long q = (long long)p - lv;
    const int rv = (q >= 0 && q < (long long)H * W) ? dr[q] : 0;
    const long long diff = llabs((long long)lv - (long long)rv);
    unsigned char c = 0;
    if (diff > 5) {
        c = lv < rv ? 1 : 2;
        out[p] = 2147483647;
    } else {
        out[p] = lv;
    }
    cls[p] = c;
}

// ---------------------------------------------------------------------------------------------- NCC, shifted-image form
// ncc() (NCC/NCC.h:117-272): for offsets 1..79 the other image is shifted by the offset (columns that would come from
// outside keep their own value), an 11x11 window clipped at the image border is correlated, and the offset with the
// strictly largest score wins (depth = offset * 3 as uchar, 0 if none beats -2).  The function's `n` is
// (end_y - start_y) * (end_x - start_x), one row and one column short of the window, and is used as the divisor of the
// means, of the three moments and once more of the score -- reproduced from exact integer window sums:
// with N taps, S1 = sum a, S2 = sum b, Q1 = sum a^2, Q2 = sum b^2, P = sum ab, ma = S1/n, mb = S2/n:
//   sum (a-ma)^2 = Q1 - 2 ma S1 + N ma^2,  sum (a-ma)(b-mb) = P - ma S2 - mb S1 + N ma mb.
__global__ void __launch_bounds__(128) ncc_shift_kernel(const unsigned char* __restrict__ left, const unsigned char* __restrict__ right,
                                                        int H, int W, int type_right, int add_constant, int max_offset, int ksize,
                                                        unsigned char* __restrict__ depth) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int sx = max(0, x - ksize), sy = max(0, y - ksize), ex = min(W - 1, x + ksize), ey = min(H - 1, y + ksize);
    const double n = (double)(ey - sy) * (double)(ex - sx);
    const double N = (double)(ey - sy + 1) * (double)(ex - sx + 1);
    double best = -2.0;
    unsigned char out = 0;
    for (int off = 1; off <= max_offset; ++off) {
        long long S1 = 0, S2 = 0, Q1 = 0, Q2 = 0, P = 0;
        for (int i = sy; i <= ey; ++i)
            for (int j = sx; j <= ex; ++j) {
                int a, b;
                if (!type_right) {  // left: a = left, b = right shifted right by off (`right += 10` saturates like cv::Mat)
                    a = left[(size_t)i * W + j];
                    const int jj = j < off ? j : j - off;
                    b = right[(size_t)i * W + jj];
                    if (add_constant) b = min(255, b + 10);
                } else {            // right: a = left shifted left by off, b = right
                    const int jj = j < W - off ? j + off : j;
                    a = left[(size_t)i * W + jj];
                    b = right[(size_t)i * W + j];
                    if (add_constant) b = min(255, b + 10);
                }
                S1 += a;
                S2 += b;
                Q1 += a * a;
                Q2 += b * b;
                P += a * b;
            }
        const double ma = (double)S1 / n, mb = (double)S2 / n;
        const double va = ((double)Q1 - 2.0 * ma * (double)S1 + N * ma * ma) / n;
        const double vb = ((double)Q2 - 2.0 * mb * (double)S2 + N * mb * mb) / n;
        const double num = ((double)P - ma * (double)S2 - mb * (double)S1 + N * ma * mb) / n;
        const double score = num / (sqrt(va) * sqrt(vb)) / n;
        if (score > best) {
            best = score;
            out = (unsigned char)(off * 3);
        }
    }
    depth[(size_t)y * W + x] = out;
}

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

smt_status_t smt_sad_crosscheck(smt_ctx* ctx, const int32_t* disp_left, const int32_t* disp_right, int32_t rows, int32_t cols,
                                int32_t* out, int32_t* occ, int32_t* n_occ, int32_t* mis, int32_t* n_mis, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp_left && disp_right && out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_sad_crosscheck: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    const int nb = (int)cdiv((long long)n, 256);
    Marshal ms(ctx, space);
    const int32_t* dl = ms.in(disp_left, n);
    const int32_t* dr = ms.in(disp_right, n);
    int32_t* dout = ms.out(out, n);
    int32_t* docc = ms.out(occ, 2 * n);
    int32_t* dmis = ms.out(mis, 2 * n);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, cls, n);
    SMT_ALLOC(ctx, counts, 2 * sizeof(int));
    {
        StageTimer t(ctx, ST_POST);
        SMT_LAUNCH(ctx, sad_crosscheck_kernel, nb, 256, 0, dl, dr, rows, cols, dout, cls.as<unsigned char>());
        SMT_TRY(k_class_map_to_lists(ctx, cls.as<unsigned char>(), cols, rows, docc, dmis, counts.as<int>()));
    }
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, counts.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_occ) *n_occ = ctx->h_pinned[0];
    if (n_mis) *n_mis = ctx->h_pinned[1];
    return ms.finish();
}

smt_status_t smt_ncc_shift(smt_ctx* ctx, const uint8_t* left, const uint8_t* right, int32_t rows, int32_t cols, int32_t right_view,
                           int32_t add_constant, uint8_t* depth, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, left && right && depth && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_ncc_shift: bad argument");
    SMT_REQUIRE(ctx, cols > 79, SMT_ERR_UNSUPPORTED, "smt_ncc_shift: the reference shifts by up to 79 columns and indexes outside narrower images");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* dl = ms.in(left, n);
    const uint8_t* dr = ms.in(right, n);
    uint8_t* dd = ms.out(depth, n);
    SMT_TRY(ms.check());
    {
        StageTimer t(ctx, ST_COST);
        SMT_LAUNCH(ctx, ncc_shift_kernel, dim3(cdiv(cols, 128), rows), 128, 0, dl, dr, rows, cols, right_view ? 1 : 0, add_constant ? 1 : 0, 79, 5,
                   dd);
    }
    return ms.finish();
}

smt_status_t smt_sad(smt_ctx* ctx, const uint8_t* left_padded, const uint8_t* right_padded, int32_t rows_padded, int32_t cols_padded,
                     int32_t D, int32_t winsize, int32_t right_view, int32_t* disp, float* costs, float* subpixel, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    const int w = winsize + 1;
    SMT_REQUIRE(ctx, left_padded && right_padded && disp && D > 0 && winsize >= 0 && rows_padded > 2 * w && cols_padded > 2 * w,
                SMT_ERR_INVALID_ARG, "smt_sad: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const int Ho = rows_padded - 2 * w, Wo = cols_padded - 2 * w;
    const size_t np = (size_t)rows_padded * cols_padded, no = (size_t)Ho * Wo;
    Marshal ms(ctx, space);
    const uint8_t* dl = ms.in(left_padded, np);
    const uint8_t* dr = ms.in(right_padded, np);
    int32_t* dd = ms.out(disp, no);
    float* dc = ms.out(costs, no * D);
    float* dsp = ms.out(subpixel, no);
    SMT_TRY(ms.check());
    SMT_REQUIRE(ctx, !(subpixel && right_view), SMT_ERR_UNSUPPORTED, "smt_sad: the sub-pixel value exists in the left view only (OptimalDisparity)");
    SMT_ALLOC(ctx, vol, no * D * sizeof(int));
    {
        StageTimer t(ctx, ST_COST);
        if (!right_view)
            SMT_TRY(k_box_sum(ctx, 0, dl, dr, rows_padded, cols_padded, w, -1, D, vol.as<int>()));
        else
            SMT_TRY(k_box_sum(ctx, 0, dr, dl, rows_padded, cols_padded, w, +1, D, vol.as<int>()));
    }
    {
        StageTimer t(ctx, ST_WTA);
        SMT_LAUNCH(ctx, sad_select_kernel, cdiv((long long)no, 128), 128, 0, vol.as<int>(), Ho, Wo, D, right_view ? 1 : 0, dd, dc, dsp);
    }
    return ms.finish();
}

smt_status_t smt_ncc(smt_ctx* ctx, const uint8_t* left, const uint8_t* right, int32_t rows, int32_t cols, int32_t win_size, int32_t D,
                     int32_t* disp, double* costs, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    const int w = win_size;
    SMT_REQUIRE(ctx, left && right && disp && D > 0 && w >= 0 && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_ncc: bad argument");
    SMT_REQUIRE(ctx, (long long)(2 * w + 1) * (2 * w + 1) * 65025 < 2147483647ll, SMT_ERR_UNSUPPORTED,
                "smt_ncc: window too large for 32-bit window sums");
    DeviceGuard dev_guard(ctx->device);
    const size_t np = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* dl = ms.in(left, np);
    const uint8_t* dr = ms.in(right, np);
    int32_t* dd = ms.out(disp, np);
    double* dcost = ms.out(costs, np * D);
    SMT_TRY(ms.check());
    SMT_CUDA(ctx, cudaMemsetAsync(dd, 0, np * sizeof(int32_t), ctx->stream));  // borders stay 0 (NCC_main.cpp:20)
    if (dcost) SMT_CUDA(ctx, cudaMemsetAsync(dcost, 0, np * D * sizeof(double), ctx->stream));  // no window there: no cost
    if (rows > 2 * w && cols > 2 * w) {
        const int Ho = rows - 2 * w, Wo = cols - 2 * w;
        SMT_ALLOC(ctx, stats, 4 * np * sizeof(int));
        SMT_ALLOC(ctx, vol, (size_t)Ho * Wo * D * sizeof(int));
        int* s = stats.as<int>();
        {
            StageTimer t(ctx, ST_COST);
            dim3 g(cdiv(cols, 128), rows);
            SMT_LAUNCH(ctx, ncc_stats_kernel, g, 128, 0, dl, rows, cols, w, s, s + np);
            SMT_LAUNCH(ctx, ncc_stats_kernel, g, 128, 0, dr, rows, cols, w, s + 2 * np, s + 3 * np);
            SMT_TRY(k_box_sum(ctx, 1, dl, dr, rows, cols, w, -1, D, vol.as<int>()));
        }
        {
            StageTimer t(ctx, ST_WTA);
            SMT_LAUNCH(ctx, ncc_select_kernel, cdiv((long long)Ho * Wo, 8), 256, 0, vol.as<int>(), s, s + np, s + 2 * np, s + 3 * np, rows,
                       cols, w, D, dd, dcost);
        }
    }
    return ms.finish();
}

smt_status_t smt_asw(smt_ctx* ctx, const uint8_t* left_padded, const uint8_t* right_padded, int32_t rows_padded, int32_t cols_padded,
                     int32_t win_size, int32_t D, double space_sigma, double color_sigma, int32_t T, int32_t right_view, float* disp,
                     float* costs, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    const int wins = win_size + 1, k = 2 * wins + 1;
    SMT_REQUIRE(ctx, left_padded && right_padded && disp && D > 0 && win_size >= 0 && rows_padded > 2 * wins && cols_padded > 2 * wins,
                SMT_ERR_INVALID_ARG, "smt_asw: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const int Ho = rows_padded - 2 * wins, Wo = cols_padded - 2 * wins;
    const size_t np = (size_t)rows_padded * cols_padded, no = (size_t)Ho * Wo;
    const size_t smem = sizeof(float) * (256 * 32 + (size_t)ASW_TX * k * k + (size_t)k * (ASW_TX + k - 1) + ASW_TX * ASW_TD) +
                        sizeof(int) * ASW_TX * ASW_TD + (size_t)k * (ASW_TX + k - 1 + D - 1 + 2 * ASW_PAD + ASW_Q);
    SMT_REQUIRE(ctx, smem <= 200 * 1024, SMT_ERR_UNSUPPORTED, "smt_asw: window / disparity range too large for the shared-memory tile");
    Marshal ms(ctx, space);
    const uint8_t* dl = ms.in(left_padded, np);
    const uint8_t* dr = ms.in(right_padded, np);
    float* dd = ms.out(disp, no);
    float* dcost = ms.out(costs, no * D);
    SMT_TRY(ms.check());
    if (dcost) SMT_CUDA(ctx, cudaMemsetAsync(dcost, 0, no * D * sizeof(float), ctx->stream));
    // masks evaluated on the host in double like the reference (ASW.h:16-47), used as float weights on the device
    std::vector<float> h((size_t)k * k + 256);
    const int c = (k - 1) / 2;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
            h[(size_t)i * k + j] = (float)std::exp(-(std::pow(j - c, 2) + std::pow(i - c, 2)) / (2 * space_sigma * space_sigma));
    for (int i = 0; i < 256; ++i) h[(size_t)k * k + i] = (float)std::exp(-(i * i) / (2 * color_sigma * color_sigma));
    SMT_ALLOC(ctx, masks, h.size() * sizeof(float));
    SMT_CUDA(ctx, cudaMemcpyAsync(masks.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // h is a local vector
    if (smem > 48 * 1024) {
        SMT_CUDA(ctx, cudaFuncSetAttribute(asw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMT_CUDA(ctx, cudaFuncSetAttribute(asw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    {
        StageTimer t(ctx, ST_COST);
        dim3 grid(cdiv(Wo, ASW_TX), Ho);
        if (right_view)
            SMT_LAUNCH(ctx, asw_kernel<true>, grid, ASW_TX * ASW_TD, smem, dr, dl, rows_padded, cols_padded, wins, D, T, masks.as<float>(),
                       masks.as<float>() + (size_t)k * k, dd, dcost);
        else
            SMT_LAUNCH(ctx, asw_kernel<false>, grid, ASW_TX * ASW_TD, smem, dl, dr, rows_padded, cols_padded, wins, D, T, masks.as<float>(),
                       masks.as<float>() + (size_t)k * k, dd, dcost);
    }
    return ms.finish();
}

smt_status_t smt_asw_crosscheck(smt_ctx* ctx, const float* disp_left, const float* disp_right, int32_t rows, int32_t cols, uint8_t* out,
                                smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp_left && disp_right && out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_asw_crosscheck: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* dl = ms.in(disp_left, n);
    const float* dr = ms.in(disp_right, n);
    uint8_t* dout = ms.out(out, n);
    SMT_TRY(ms.check());
    SMT_TRY(k_asw_crosscheck(ctx, dl, dr, rows, cols, dout));
    return ms.finish();
}

}  // extern "C"
