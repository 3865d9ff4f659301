// SMT_SUM_FAST cross aggregation, fixed-point streaming form (the hot path of P_full).
//
// Same walk as cross_agg_seq.cu -- one warp walks one whole line, lane = disparity, a pass along the line is
// out[j] = E[j+hi(j)+1] - E[j-lo(j)] with E the exclusive prefix of the line (cross_aggregator.cpp:353-389), the last 72
// prefix entries live in a lane-private shared-memory ring -- but the prefix is a 32-bit WRAP-AROUND FIXED-POINT sum:
//   * every tap is quantised once to the quantum q (2^32 q >= 128 x the power of two above the largest cost) and added
//     as an integer; integer prefixes may overflow freely, because E[b] - E[a] is exact modulo 2^32 and a window of
//     <= 69 taps stays below 2^32 quanta.  A ring entry is 4 bytes (compensated float pair: 8), a prefix step is one IADD
//     (4 FADD), a window sum one IADD (3 FADD): half the shared-memory wavefronts and half the instructions per cell;
//   * the dividing ("primed") pass of an iteration sums the raw sums of the other direction (taps < 69 S, windows up
//     to 4761 S): it keeps two prefixes, a fine one (q; exact for pixels whose support count is < 128) and a coarse
//     one (64 q, the six dropped bits carried into the next tap, so a window sum is off by < 1 coarse quantum whatever
//     its length), in two ring planes, and each output reads the plane its own support count selects;
//   * that makes the ring small enough to run a dividing pass and the following non-dividing pass of the SAME direction
//     (the reference's order H,V | V,H | H,V | V,H contains V,V / H,H / V,V) in ONE walk at 8 warps per SM: the
//     intermediate volume of the pair never exists, 8 passes are 5 sweeps over the volume.
// Accuracy.  A window sum is exact up to the quantisation of its taps and rounded to float32 once.  Taps are
// non-negative, so the relative error of a window sum is at most the largest relative quantisation error among its
// non-zero taps, (q/2) / v; for the coarse plane it is < 64 q / (sup x mean) <= (q/2) / mean.  Every value that is about
// to become a tap (pass inputs, the means a dividing pass emits) is therefore checked: zero, or at least `vmin` = 2^19
// quanta, i.e. a relative error <= 2^-20 = 9.5e-7 per quantisation and < 1e-5 over the eight passes of an aggregation
// (for AD-Census costs 2^19 quanta are 0.031, just below the smallest non-zero cost 1 - exp(-1/30)).  A warp that meets a
// smaller value appends its (line, disparity chunk) to the sweep's redo list; the float-pair passes of
// cross_agg_seq.cu, queued behind every fixed-point launch, then redo exactly those lines (on the benchmark frames: a
// few per cent, the lines through low-cost texture) -- or the whole sweep, and all later ones, once more than a quarter
// of the lines asked for it (FxState::escalated), which is what volumes with a wide dynamic range such as CBLSM's integer
// AD costs end up doing.  Measured against the compiled reference: DESIGN.md 4 / tests/test_gpu_baseline_configs.py.
#include "common.cuh"

namespace {

constexpr int FX_CH = 24;     // positions per chunk
constexpr int FX_R = 72;      // ring slots = 3 chunks (span needed: 71)
constexpr int FX_LOOK = 35;   // output j is computable once E[j+35] is stored
#ifndef SMT_FX_DLAG
#define SMT_FX_DLAG 4
#endif
#ifndef SMT_FX_RLAG
#define SMT_FX_RLAG 4
#endif
constexpr int FX_DLAG = SMT_FX_DLAG;  // descriptor read this many positions ahead of its ring reads
constexpr int FX_RLAG = SMT_FX_RLAG;  // ring reads consumed this many positions after they are issued
constexpr int FX_PLANE = FX_R * 128;  // bytes of one ring plane (72 slots x 32 lanes x 4 B)
#ifndef SMT_FX_L2PF
#define SMT_FX_L2PF 2         // chunks between the register loads and the L2 prefetch running ahead of them
#endif

// correctly rounded a / b given y = RN(1/b) (Markstein); b = support count times a power of two
__device__ __forceinline__ float fx_div(float a, float b, float y) {
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

// window descriptor of one output position: byte offsets of E[j-lo] and E[j+hi+1] in the lane's ring column (the plane
// of a dividing pass included), and for a dividing pass  bq = count / quantum, yq = RN(1/count) x quantum  (exact
// quotient)
struct __align__(16) FxDesc {
    unsigned offa, offb;
    float bq, yq;
};
// the fused pair reads ONE 16-byte record per step: byte offsets of A's window of item k (plane included), A's quantum
// ratio / count, and the two 16-bit byte offsets of B's window of item k - FX2_S3.  (An 8-byte record with four slot
// indices, one LDS.64 and byte extraction measured 2.5 % slower: the extra PRMT / IMAD cost more than the narrower load.)
struct __align__(16) FxRec {
    unsigned offa, offb;
    float rq;
    unsigned offs_b;  // B: offa | offb << 16
};

struct FxPre {  // raw {left,right,top,bottom} arm word and support count of this lane's output position in the NEXT chunk
    unsigned armw;
    int cnt;
};

__device__ __forceinline__ int fx_slot(int tbm, int rel) {  // ring slot of position tb + rel, rel in [-72, 71]
    int r = tbm + rel + FX_R;
    if (r >= FX_R) r -= FX_R;
    if (r >= FX_R) r -= FX_R;
    return r;
}

// a warp whose line left the domain of the error bound asks for the float-pair passes to redo it
__device__ __forceinline__ void fx_request_redo(int* state, int task) {
    const int n = atomicAdd(state + FX_ST_COUNT, 1);
    atomicAdd(state + FX_ST_TOTAL, 1);
    state[FX_ST_LIST + n] = task;
    if (n >= state[FX_ST_LIMIT]) atomicOr(state + FX_ST_ESCALATED, 1);
}

// prefix state of one stage
struct FxAcc {
    unsigned fine = 0u, coarse = 0u, carry = 0u;
};
template <bool PRIMED>
__device__ __forceinline__ void fx_push(FxAcc& a, unsigned char* slot, unsigned tq) {
    *reinterpret_cast<unsigned*>(slot) = a.fine;
    a.fine += tq;
    if (PRIMED) {
        *reinterpret_cast<unsigned*>(slot + FX_PLANE) = a.coarse;
        const unsigned c = a.carry + tq;
        a.coarse += c >> 6;
        a.carry = c & 63u;
    }
}

// window of output position j (arms from `aw`) in a ring whose slot of position p is (p + shift) % 72
__device__ __forceinline__ void fx_window(int j, unsigned aw, int len, int tb, int tbm, int shift, unsigned& offa, unsigned& offb) {
    const int lo = (int)(aw & 0xffu), hi = (int)(aw >> 8);
    // arms stop at the image border; the clamps only act on the halo rows of a row band, whose results are never used
    const int a = max(j - lo, 0), b = min(j + hi + 1, len);
    offa = (unsigned)fx_slot(tbm, a + shift - tb) * 128u;
    offb = (unsigned)fx_slot(tbm, b + shift - tb) * 128u;
}

__device__ __forceinline__ unsigned fx_quant(float v, float inv_q) { return __float2uint_rn(__fmul_rn(v, inv_q)); }
// a value about to become a tap must be 0 or >= vmin quanta (vm1 = vmin - 1)
__device__ __forceinline__ unsigned fx_small(unsigned tq, unsigned vm1) { return (unsigned)((tq - 1u) < vm1); }

template <int CSTEP>
__device__ __forceinline__ void fx_load_chunk(float (&vload)[FX_CH], const float* __restrict__ src_next, long long step, int dep, bool edge,
                                              int tb, int len, bool act, int lane) {
    const long long step_b = step * 4;
    const char* p = reinterpret_cast<const char*>(src_next + dep);
#pragma unroll
    for (int k = 0; k < FX_CH; ++k) {
        if (edge)
            vload[k] = (tb + FX_CH + k < len && act) ? __ldcs(reinterpret_cast<const float*>(p)) : 0.f;
        else
            vload[k] = __ldcs(reinterpret_cast<const float*>(p));
        p += step_b;
    }
#if SMT_FX_L2PF > 0
    {   // DRAM -> L2 runs further ahead than the one load scoreboard lets registers: lane l asks for the row of position
        // tb + (1 + L2PF) * 24 + l with one instruction per chunk
        const int qn = tb + (1 + SMT_FX_L2PF) * FX_CH + lane;
        if (lane < FX_CH && qn < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(src_next + dep + step * (SMT_FX_L2PF * FX_CH + lane)));
    }
#endif
}

// =========================================================================================== single pass (one stage)
// PRIMED: the dividing pass of an iteration (input: raw sums of the other direction; output: means, exact quotient);
// otherwise the non-dividing pass (input: costs or means; output: raw sums).
template <bool PRIMED, bool EDGE, int CSTEP>
__device__ __forceinline__ void fx1_chunk(FxAcc& acc, unsigned& bad, FxPre& pre, float (&vuse)[FX_CH], float (&vload)[FX_CH], int tb,
                                          int tbm, int len, bool act, const float* __restrict__ src_next, float* __restrict__ dst_first,
                                          long long step_rt, long long mbase, long long mstride, int horiz,
                                          const unsigned* __restrict__ arms, const uint16_t* __restrict__ sup, unsigned char* ring_col,
                                          FxDesc* desc, int lane, float inv_q, float q, unsigned vm1, int zero, int st_lo, int st_hi) {
    const long long step = CSTEP ? (long long)CSTEP : step_rt;
    const long long step_b = step * 4;
    // ---- (1a) descriptors of this chunk's outputs (position tb - 35 + lane) from the arm word / count fetched a chunk ago
    {
        const int j = tb - FX_LOOK + lane;
        FxDesc ds{0u, 0u, 1.f, 1.f};
        if (lane < FX_CH && (!EDGE || (j >= 0 && j < len))) {
            const unsigned aw = horiz ? (pre.armw & 0xffffu) : (pre.armw >> 16);
            fx_window(j, aw, len, tb, tbm, 0, ds.offa, ds.offb);
            if (PRIMED) {
                const bool coarse = pre.cnt >= 128;
                const float c = (float)pre.cnt;
                if (coarse) {
                    ds.offa += FX_PLANE;
                    ds.offb += FX_PLANE;
                }
                ds.bq = __fmul_rn(c, coarse ? inv_q * 0.015625f : inv_q);
                ds.yq = __fmul_rn(__frcp_rn(c), coarse ? q * 64.f : q);
            }
        }
        __syncwarp();  // the previous chunk's descriptor reads are done
        if (lane < FX_CH) desc[lane] = ds;
        __syncwarp();
    }
    unsigned char* wbase = ring_col + tbm * 128;   // slot of E[tb]
    char* dp = reinterpret_cast<char*>(dst_first);  // output position tb - 35
    // ---- (1b) quantise the chunk (values loaded a whole chunk ago) and take the first prefix step
    unsigned tq[FX_CH];
#pragma unroll
    for (int k = 0; k < FX_CH; ++k) {
        tq[k] = fx_quant(vuse[k], inv_q);
        if (!PRIMED) bad |= fx_small(tq[k], vm1);  // dividing passes take the raw sums a guarded pass has produced
    }
    fx_push<PRIMED>(acc, wbase, tq[0]);
    // ---- (2) loads of the next chunk; the address depends on (1b) so that they cannot be hoisted above it: all global
    // loads share one scoreboard, and the first consumer of ANY loaded register waits for ALL loads in flight
    const int dep = (int)acc.fine & zero;
    fx_load_chunk<CSTEP>(vload, src_next, step, dep, EDGE, tb, len, act, lane);
    {
        const int jn = tb - FX_LOOK + lane + FX_CH;
        pre.armw = 0u;
        pre.cnt = 1;
        if (lane < FX_CH && jn >= 0 && jn < len) {
            const long long pix = mbase + mstride * jn + dep;
            pre.armw = __ldg(arms + pix);
            if (PRIMED) pre.cnt = (int)__ldg(sup + pix);
        }
    }
    // ---- (3) remaining prefix steps and the outputs, software-pipelined over the chunk
    {
        FxDesc dsc[FX_CH];
        unsigned ea[FX_CH], eb[FX_CH];
#pragma unroll
        for (int u = 0; u < FX_DLAG; ++u) dsc[u] = desc[u];
#pragma unroll
        for (int k = 0; k < FX_CH + FX_RLAG; ++k) {
            if (k < FX_CH) {
                if (k > 0) fx_push<PRIMED>(acc, wbase + k * 128, tq[k]);
                ea[k] = *reinterpret_cast<const unsigned*>(ring_col + dsc[k].offa);
                eb[k] = *reinterpret_cast<const unsigned*>(ring_col + dsc[k].offb);
                if (k + FX_DLAG < FX_CH) dsc[k + FX_DLAG] = desc[k + FX_DLAG];
            }
            if (k >= FX_RLAG) {
                const int kk = k - FX_RLAG;
                const bool inside = !EDGE || (tb - FX_LOOK + kk >= st_lo && tb - FX_LOOK + kk < st_hi);  // warp-uniform
                float r = __uint2float_rn(eb[kk] - ea[kk]);
                if (PRIMED) {
                    r = fx_div(r, dsc[kk].bq, dsc[kk].yq);
                    // The taps of a dividing pass are not checked one by one (a raw sum of a redone line may be small):
                    // the MEAN is.  The error of the window sum is at most (taps) x q/2 and the sum is count x mean with
                    // count >= taps, so mean >= vmin bounds the relative error by 2^-20 like a per-tap check would.
                    bad |= fx_small(fx_quant(r, inv_q), vm1);
                } else {
                    r = __fmul_rn(r, q);
                }
                if (inside && (!EDGE || act)) __stcs(reinterpret_cast<float*>(dp), r);
                dp += step_b;
            }
        }
    }
}

constexpr int FX1_WARPS = 4;
template <bool PRIMED>
struct Fx1Cfg {
    static constexpr int SMEM_PER_WARP = FX_PLANE * (PRIMED ? 2 : 1);
};

template <bool PRIMED, int CSTEP>
__global__ void __launch_bounds__(FX1_WARPS * 32, PRIMED ? 3 : 4)
    agg_fx1_kernel(const float* __restrict__ in, float* __restrict__ out, const unsigned* __restrict__ arms,
                   const uint16_t* __restrict__ sup, int H, int W, int D, int horiz, int tmaps, float inv_q, float q, unsigned vm1,
                   int* __restrict__ state, int zero, int st_lo, int st_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FxDesc s_desc_all[FX1_WARPS][FX_CH];
    if (*reinterpret_cast<volatile int*>(state + FX_ST_ESCALATED) != 0) return;  // the float-pair passes do the whole sweep
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* ring_col = smem_raw + (size_t)warp * Fx1Cfg<PRIMED>::SMEM_PER_WARP + lane * 4;
    FxDesc* s_desc = s_desc_all[warp];
    const int nchunk = (D + 31) >> 5;
    const long long wg = (long long)blockIdx.x * FX1_WARPS + warp;
    const int line = (int)(wg / nchunk), chunk = (int)(wg - (long long)line * nchunk);
    const int len = horiz ? W : H;
    if (line >= (horiz ? H : W)) return;
    const int d = chunk * 32 + lane;
    const bool act = d < D;
    const bool full = chunk * 32 + 32 <= D;
    const long long pstride = horiz ? 1 : (long long)W;
    const long long pbase = horiz ? (long long)line * W : (long long)line;
    const long long step = CSTEP ? (long long)CSTEP : pstride * D;
    const float* src = in + pbase * D + (act ? d : 0);
    float* dst = out + pbase * D + (act ? d : 0);
    const long long mstride = (!horiz && tmaps) ? 1 : pstride;
    const long long mbase = (!horiz && tmaps) ? (long long)line * H : pbase;

    FxAcc acc;
    unsigned bad = 0u;
    float va[FX_CH], vb[FX_CH];
#pragma unroll
    for (int k = 0; k < FX_CH; ++k) va[k] = (k < len && act) ? __ldcs(src + step * k) : 0.f;
    FxPre pre{0u, 1};
    const int nouter = (len + FX_LOOK + FX_CH - 1) / FX_CH;
    int tbm = 0;
#define SMT_FX1_STEP(USE, LOAD)                                                                                              \
    if (o < nouter) {                                                                                                         \
        const int tb = o * FX_CH;                                                                                             \
        const float* src_next = src + step * (tb + FX_CH);                                                                    \
        float* dst_first = dst + step * (tb - FX_LOOK);                                                                       \
        /* interior: every load and every output of the chunk is inside the line and inside the stored range */            \
        if (full && tb - FX_LOOK >= st_lo && tb - FX_LOOK + FX_CH <= st_hi && tb + 2 * FX_CH <= len)                          \
            fx1_chunk<PRIMED, false, CSTEP>(acc, bad, pre, USE, LOAD, tb, tbm, len, act, src_next, dst_first, step, mbase,    \
                                            mstride, horiz, arms, sup, ring_col, s_desc, lane, inv_q, q, vm1, zero, st_lo,    \
                                            st_hi);                                                                           \
        else                                                                                                                  \
            fx1_chunk<PRIMED, true, CSTEP>(acc, bad, pre, USE, LOAD, tb, tbm, len, act, src_next, dst_first, step, mbase,     \
                                           mstride, horiz, arms, sup, ring_col, s_desc, lane, inv_q, q, vm1, zero, st_lo,     \
                                           st_hi);                                                                            \
        tbm = tbm == FX_R - FX_CH ? 0 : tbm + FX_CH;                                                                          \
        ++o;                                                                                                                  \
    }
    for (int o = 0; o < nouter;) {
        SMT_FX1_STEP(va, vb)
        SMT_FX1_STEP(vb, va)
    }
#undef SMT_FX1_STEP
    if (__any_sync(0xffffffffu, bad != 0u) && lane == 0) fx_request_redo(state, (int)wg);
}

// ===================================================================================== fused pair (dividing pass A, then B)
// Four stages run over the positions of a chunk, each a fixed number of positions behind the previous one, so that no
// shared-memory access of an iteration waits for arithmetic of the same iteration (a warp issues in order, and only two
// warps share a scheduler here).  Item i of a chunk is A's output position tb+i-35, which is B's tap position.
//   S1  i = k            A: store the prefix of position tb+k, add its tap; issue the two ring reads of A's output i
//   S2  i = k-RLAG       A's output: window sum -> mean in quanta (x plane ratio / count) = B's tap
//   S3  i = k-RLAG-1     B: store the prefix of position tb+i-35, add the tap S2 produced one iteration ago; issue the
//                        ring reads of B's output tb+i-70.  B's ring slot of position p is (p+35) % 72, i.e. tbm+i:
//                        both rings advance with compile-time offsets
//   S4  i = k-2RLAG-1    B's output: window sum x quantum -> store
// Items with a negative index belong to the previous chunk; their in-flight values cross the chunk boundary in registers
// (FxCarry).  B's window descriptors were prepared together with A's one or two chunks earlier and wait in a three-chunk
// descriptor ring.
#ifndef SMT_FX2_ONEBODY
#define SMT_FX2_ONEBODY 1
#endif
constexpr int FX2_WARPS = 4;
constexpr int FX2_S3 = FX_RLAG + 1, FX2_S4 = 2 * FX_RLAG + 1;  // stage lags in positions
constexpr int FX2_SMEM_PER_WARP = 3 * FX_PLANE + 3 * FX_CH * (int)sizeof(unsigned);

struct FxCarry {
    unsigned eaA[FX_RLAG], ebA[FX_RLAG];
    float rq[FX_RLAG];
    unsigned tqb;
    unsigned eaB[FX_RLAG], ebB[FX_RLAG];
};

template <bool EDGE, int CSTEP>
__device__ __forceinline__ void fx2_chunk(FxAcc& accA, FxAcc& accB, FxCarry& cy, unsigned& bad, FxPre& pre, float (&vuse)[FX_CH],
                                          float (&vload)[FX_CH], int tb, int tbm, int o3, int len, bool act,
                                          const float* __restrict__ src_next, float* __restrict__ dst_first, long long step_rt,
                                          long long mbase, long long mstride, int horiz, const unsigned* __restrict__ arms,
                                          const uint16_t* __restrict__ sup, unsigned char* ringA, unsigned char* ringB, FxRec* descA,
                                          unsigned* descB, int lane, float inv_q, float q, unsigned vm1, int zero) {
    const long long step = CSTEP ? (long long)CSTEP : step_rt;
    const long long step_b = step * 4;
    // B's window slots wait in a three-chunk ring: this chunk writes row o3, reads rows (o3+2)%3 (one chunk ago) and
    // (o3+1)%3 (two chunks ago)
    unsigned* const dB_w = descB + o3 * FX_CH;
    const unsigned* const dB_1 = descB + (o3 == 0 ? 2 : o3 - 1) * FX_CH;
    const unsigned* const dB_2 = descB + (o3 == 2 ? 0 : o3 + 1) * FX_CH;
    {
        const int j = tb - FX_LOOK + lane;
        FxRec rec{0u, 0u, 0.f, 0u};
        unsigned db = 0u;
        if (lane < FX_CH && (!EDGE || (j >= 0 && j < len))) {
            const unsigned aw = horiz ? (pre.armw & 0xffffu) : (pre.armw >> 16);
            unsigned oa, ob, pa, pb;
            fx_window(j, aw, len, tb, tbm, 0, oa, ob);
            fx_window(j, aw, len, tb, tbm, FX_LOOK, pa, pb);
            const bool coarse = pre.cnt >= 128;
            const unsigned plane = coarse ? (unsigned)FX_PLANE : 0u;
            rec.offa = oa + plane;
            rec.offb = ob + plane;
            db = pa | (pb << 16);
            rec.rq = __fmul_rn(__frcp_rn((float)pre.cnt), coarse ? 64.f : 1.f);  // window sum in its plane's quanta -> mean in quanta
        }
        __syncwarp();
        // one record per step: A's window of item `lane` and B's window of item lane - FX2_S3, prepared one or two chunks
        // ago: row o-1 at i-11 (i >= 11) or row o-2 at i+13
        if (lane < FX_CH) {
            dB_w[lane] = db;
            const int i = lane - FX2_S3;
            rec.offs_b = (i < 11) ? dB_2[i + 13] : dB_1[i - 11];
            descA[lane] = rec;
        }
        __syncwarp();
    }
    unsigned char* wA = ringA + tbm * 128;
    unsigned char* wB = ringB + tbm * 128;                                       // B slots of this chunk's items i >= 0
    unsigned char* wBp = ringB + (tbm == 0 ? FX_R - FX_CH : tbm - FX_CH) * 128;  // ... of the previous chunk's (i < 0)
    char* dp = reinterpret_cast<char*>(dst_first);                               // output position tb - 70 - FX2_S4
    unsigned tq[FX_CH];
#pragma unroll
    for (int k = 0; k < FX_CH; ++k) tq[k] = fx_quant(vuse[k], inv_q);
    fx_push<true>(accA, wA, tq[0]);
    const int dep = (int)accA.fine & zero;
    fx_load_chunk<CSTEP>(vload, src_next, step, dep, EDGE, tb, len, act, lane);
    {
        const int jn = tb - FX_LOOK + lane + FX_CH;
        pre.armw = 0u;
        pre.cnt = 1;
        if (lane < FX_CH && jn >= 0 && jn < len) {
            const long long pix = mbase + mstride * jn + dep;
            pre.armw = __ldg(arms + pix);
            pre.cnt = (int)__ldg(sup + pix);
        }
    }
    {
        // index = item + lag, so that the previous chunk's in-flight items sit at the front
        unsigned eaA[FX_RLAG + FX_CH], ebA[FX_RLAG + FX_CH];  // item i at [i + RLAG]
        float rq[FX_RLAG + FX_CH];
        unsigned eaB[FX_RLAG + FX_CH], ebB[FX_RLAG + FX_CH];  // item i at [i + FX2_S4]
        FxRec dA[FX_CH];                                      // step k: A's item k, B's item k - FX2_S3
#pragma unroll
        for (int u = 0; u < FX_RLAG; ++u) {
            eaA[u] = cy.eaA[u];
            ebA[u] = cy.ebA[u];
            rq[u] = cy.rq[u];
            eaB[u] = cy.eaB[u];
            ebB[u] = cy.ebB[u];
        }
        unsigned tqb = cy.tqb;
#pragma unroll
        for (int u = 0; u < FX_DLAG; ++u) dA[u] = descA[u];
#pragma unroll
        for (int k = 0; k < FX_CH; ++k) {
            // ---- the shared-memory traffic of the iteration first: none of it waits for this iteration's arithmetic
            if (k > 0) fx_push<true>(accA, wA + k * 128, tq[k]);  // S1
            {
                const int i = k - FX2_S3;                          // S3: tap = the mean S2 produced one iteration ago
                fx_push<false>(accB, i >= 0 ? wB + i * 128 : wBp + (FX_CH + i) * 128, tqb);
            }
            eaA[FX_RLAG + k] = *reinterpret_cast<const unsigned*>(ringA + dA[k].offa);
            ebA[FX_RLAG + k] = *reinterpret_cast<const unsigned*>(ringA + dA[k].offb);
            rq[FX_RLAG + k] = dA[k].rq;
            eaB[FX_RLAG + k] = *reinterpret_cast<const unsigned*>(ringB + (dA[k].offs_b & 0xffffu));
            ebB[FX_RLAG + k] = *reinterpret_cast<const unsigned*>(ringB + (dA[k].offs_b >> 16));
            if (k + FX_DLAG < FX_CH) dA[k + FX_DLAG] = descA[k + FX_DLAG];
            // ---- S2: A's output of item k - RLAG
            {
                const float mean_q = __fmul_rn(__uint2float_rn(ebA[k] - eaA[k]), rq[k]);
                tqb = __float2uint_rn(mean_q);
                bad |= fx_small(tqb, vm1);
            }
            // ---- S4: B's output of item k - FX2_S4
            {
                const bool inside = !EDGE || (tb - 2 * FX_LOOK - FX2_S4 + k >= 0 && tb - 2 * FX_LOOK - FX2_S4 + k < len);
                const float r = __fmul_rn(__uint2float_rn(ebB[k] - eaB[k]), q);
                if (inside && (!EDGE || act)) __stcs(reinterpret_cast<float*>(dp), r);
                dp += step_b;
            }
        }
#pragma unroll
        for (int u = 0; u < FX_RLAG; ++u) {
            cy.eaA[u] = eaA[FX_CH + u];
            cy.ebA[u] = ebA[FX_CH + u];
            cy.rq[u] = rq[FX_CH + u];
            cy.eaB[u] = eaB[FX_CH + u];
            cy.ebB[u] = ebB[FX_CH + u];
        }
        cy.tqb = tqb;
    }
}

template <int CSTEP>
__global__ void __launch_bounds__(FX2_WARPS * 32, 2)
    agg_fx2_kernel(const float* __restrict__ in, float* __restrict__ out, const unsigned* __restrict__ arms,
                   const uint16_t* __restrict__ sup, int H, int W, int D, int horiz, int tmaps, float inv_q, float q, unsigned vm1,
                   int* __restrict__ state, int zero) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FxRec s_descA_all[FX2_WARPS][FX_CH];
    if (*reinterpret_cast<volatile int*>(state + FX_ST_ESCALATED) != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smem_raw + (size_t)warp * FX2_SMEM_PER_WARP;
    unsigned char* ringA = base + lane * 4;                 // two planes
    unsigned char* ringB = base + 2 * FX_PLANE + lane * 4;  // one plane
    unsigned* descB = reinterpret_cast<unsigned*>(base + 3 * FX_PLANE);
    FxRec* descA = s_descA_all[warp];
    const int nchunk = (D + 31) >> 5;
    const long long wg = (long long)blockIdx.x * FX2_WARPS + warp;
    const int line = (int)(wg / nchunk), chunk = (int)(wg - (long long)line * nchunk);
    const int len = horiz ? W : H;
    if (line >= (horiz ? H : W)) return;
    const int d = chunk * 32 + lane;
    const bool act = d < D;
    const bool full = chunk * 32 + 32 <= D;
    const long long pstride = horiz ? 1 : (long long)W;
    const long long pbase = horiz ? (long long)line * W : (long long)line;
    const long long step = CSTEP ? (long long)CSTEP : pstride * D;
    const float* src = in + pbase * D + (act ? d : 0);
    float* dst = out + pbase * D + (act ? d : 0);
    const long long mstride = (!horiz && tmaps) ? 1 : pstride;
    const long long mbase = (!horiz && tmaps) ? (long long)line * H : pbase;

    // B's outputs of the first chunks lie before the line: their descriptors must read as "empty window"; slot 0 of both
    // rings is read through empty descriptors before anything is stored there
    for (int k = lane; k < 3 * FX_CH; k += 32) descB[k] = 0;
    *reinterpret_cast<unsigned*>(ringA) = 0u;
    *reinterpret_cast<unsigned*>(ringB) = 0u;
    __syncwarp();

    FxAcc accA, accB;
    FxCarry cy;
#pragma unroll
    for (int u = 0; u < FX_RLAG; ++u) {
        cy.eaA[u] = cy.ebA[u] = cy.eaB[u] = cy.ebB[u] = 0u;
        cy.rq[u] = 0.f;
    }
    cy.tqb = 0u;
    unsigned bad = 0u;
    float va[FX_CH], vb[FX_CH];
#pragma unroll
    for (int k = 0; k < FX_CH; ++k) va[k] = (k < len && act) ? __ldcs(src + step * k) : 0.f;
    FxPre pre{0u, 1};
    const int nouter = (len + 2 * FX_LOOK + FX2_S4 + FX_CH - 1) / FX_CH;
    int tbm = 0, o3 = 0;
#define SMT_FX2_STEP(USE, LOAD)                                                                                              \
    if (o < nouter) {                                                                                                         \
        const int tb = o * FX_CH;                                                                                             \
        const float* src_next = src + step * (tb + FX_CH);                                                                    \
        float* dst_first = dst + step * (tb - 2 * FX_LOOK - FX2_S4);                                                          \
        if (full && tb - 2 * FX_LOOK - FX2_S4 >= 0 && tb + 2 * FX_CH <= len)                                                  \
            fx2_chunk<false, CSTEP>(accA, accB, cy, bad, pre, USE, LOAD, tb, tbm, o3, len, act, src_next, dst_first, step,    \
                                    mbase, mstride, horiz, arms, sup, ringA, ringB, descA, descB, lane, inv_q, q, vm1, zero); \
        else                                                                                                                  \
            fx2_chunk<true, CSTEP>(accA, accB, cy, bad, pre, USE, LOAD, tb, tbm, o3, len, act, src_next, dst_first, step,     \
                                   mbase, mstride, horiz, arms, sup, ringA, ringB, descA, descB, lane, inv_q, q, vm1, zero);  \
        tbm = tbm == FX_R - FX_CH ? 0 : tbm + FX_CH;                                                                          \
        o3 = o3 == 2 ? 0 : o3 + 1;                                                                                            \
        ++o;                                                                                                                  \
    }
#if SMT_FX2_ONEBODY
    // one copy of the chunk body instead of two with the buffers swapped: 24 register moves per chunk buy half the code
    // (the body is ~1000 instructions and only two warps share a scheduler's instruction cache)
    for (int o = 0; o < nouter;) {
        SMT_FX2_STEP(va, vb)
#pragma unroll
        for (int k = 0; k < FX_CH; ++k) va[k] = vb[k];
    }
#else
    for (int o = 0; o < nouter;) {
        SMT_FX2_STEP(va, vb)
        SMT_FX2_STEP(vb, va)
    }
#endif
#undef SMT_FX2_STEP
    if (__any_sync(0xffffffffu, bad != 0u) && lane == 0) fx_request_redo(state, (int)wg);
}

template <typename K>
static smt_status_t fx_set_smem(smt_ctx* ctx, K kernel, bool& done, size_t smem) {
    if (!done) {
        SMT_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMT_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        done = true;
    }
    return SMT_OK;
}

constexpr unsigned FX_VMIN = 1u << 19;  // smallest non-zero tap the fixed-point route admits, in quanta

template <bool PRIMED, int CSTEP>
static smt_status_t launch_fx1(smt_ctx* ctx, const float* in, float* out, const unsigned* arms, const uint16_t* sup, int H, int W, int D,
                               int horiz, int tmaps, float inv_q, float q, int* state, int st_lo, int st_hi) {
    constexpr size_t smem = (size_t)FX1_WARPS * Fx1Cfg<PRIMED>::SMEM_PER_WARP;
    static bool done[64] = {};
    SMT_TRY(fx_set_smem(ctx, agg_fx1_kernel<PRIMED, CSTEP>, done[ctx->device & 63], smem));
    const long long nwarps = (long long)(horiz ? H : W) * ((D + 31) / 32);
    SMT_LAUNCH(ctx, (agg_fx1_kernel<PRIMED, CSTEP>), cdiv(nwarps, FX1_WARPS), FX1_WARPS * 32, smem, in, out, arms, sup, H, W, D, horiz, tmaps,
               inv_q, q, FX_VMIN - 1u, state, 0, st_lo, st_hi);
    return SMT_OK;
}

template <int CSTEP>
static smt_status_t launch_fx2(smt_ctx* ctx, const float* in, float* out, const unsigned* arms, const uint16_t* sup, int H, int W, int D,
                               int horiz, int tmaps, float inv_q, float q, int* state) {
    constexpr size_t smem = (size_t)FX2_WARPS * FX2_SMEM_PER_WARP;
    static bool done[64] = {};
    SMT_TRY(fx_set_smem(ctx, agg_fx2_kernel<CSTEP>, done[ctx->device & 63], smem));
    const long long nwarps = (long long)(horiz ? H : W) * ((D + 31) / 32);
    SMT_LAUNCH(ctx, (agg_fx2_kernel<CSTEP>), cdiv(nwarps, FX2_WARPS), FX2_WARPS * 32, smem, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q,
               q, FX_VMIN - 1u, state, 0);
    return SMT_OK;
}

// position steps (in floats) with a compile-time instantiation: loads and stores of a chunk then use immediate offsets
// from few base registers.  256: rows of a 256-disparity volume; 491520: columns of a 1920 x 256 volume (the benchmark).
constexpr int FX_STEP_H256 = 256, FX_STEP_V1080P = 1920 * 256;

}  // namespace

#ifdef SMT_DEV_VARIANTS
#include <cstdlib>
smt_status_t k_agg_fx_pass_tma(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, int H, int W, int D, int horiz, int tmaps,
                               float inv_q, float q, int* state);
#endif

// quantum of the fixed-point prefixes for costs below `vmax`: q = S * 2^-25 with S the power of two >= vmax, so that 69
// taps stay below 2^32 quanta and dyadic inputs (integer-valued volumes) are represented exactly.  Returns false when the
// fixed-point form does not apply (unknown sign / range, or a range whose quantum is not a normal float).
bool fx_quantum(float vmax, float* inv_q, float* q) {
    if (!(vmax > 0.f) || !(vmax < 1e18f) || vmax < 1e-18f) return false;
    float s = 1.f;
    while (s < vmax) s *= 2.f;
    while (s * 0.5f >= vmax) s *= 0.5f;
    *q = s * (1.f / 33554432.f);
    *inv_q = 33554432.f / s;
    return true;
}

// one pass; sup != nullptr makes it the dividing pass.  `state`: the sweep's redo state (FX_ST_*), count reset by the caller.
// st_lo, st_hi: only outputs at positions [st_lo, st_hi) of each line are stored (row bands: a vertical pass must leave the
// halo rows of its output alone, the neighbours write those); st_hi <= 0 means the whole line.
smt_status_t k_agg_fx_pass(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                           int horiz, int tmaps, float inv_q, float q, int* state, int st_lo, int st_hi) {
    const unsigned* arms = reinterpret_cast<const unsigned*>(arms8);
    const long long step = horiz ? D : (long long)W * D;
    if (st_hi <= 0) {
        st_lo = 0;
        st_hi = horiz ? W : H;
    }
#define SMT_FX1_DISPATCH(PRIMED_)                                                                                                   \
    if (step == FX_STEP_H256) return launch_fx1<PRIMED_, FX_STEP_H256>(ctx, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q, q, state, st_lo, st_hi);   \
    if (step == FX_STEP_V1080P) return launch_fx1<PRIMED_, FX_STEP_V1080P>(ctx, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q, q, state, st_lo, st_hi); \
    return launch_fx1<PRIMED_, 0>(ctx, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q, q, state, st_lo, st_hi);
    if (sup) {
        SMT_FX1_DISPATCH(true)
    }
#ifdef SMT_DEV_VARIANTS
    // A/B: the bulk-copy staged form of the non-dividing pass (dev/fx1_tma.cuh), full lines only
    static const bool use_tma = getenv("SMT_FX1_TMA") != nullptr;
    if (use_tma && D % 32 == 0 && st_lo == 0 && st_hi == (horiz ? W : H))
        return k_agg_fx_pass_tma(ctx, in, out, arms8, H, W, D, horiz, tmaps, inv_q, q, state);
#endif
    SMT_FX1_DISPATCH(false)
#undef SMT_FX1_DISPATCH
}

// dividing pass (by sup) followed by the non-dividing pass of the next iteration along the same direction, in one walk
smt_status_t k_agg_fx_pair(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                           int horiz, int tmaps, float inv_q, float q, int* state) {
    const unsigned* arms = reinterpret_cast<const unsigned*>(arms8);
    const long long step = horiz ? D : (long long)W * D;
    if (step == FX_STEP_H256) return launch_fx2<FX_STEP_H256>(ctx, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q, q, state);
    if (step == FX_STEP_V1080P) return launch_fx2<FX_STEP_V1080P>(ctx, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q, q, state);
    return launch_fx2<0>(ctx, in, out, arms, sup, H, W, D, horiz, tmaps, inv_q, q, state);
}

#ifdef SMT_DEV_VARIANTS
#include "dev/fx1_tma.cuh"
#endif
