// SMT_SUM_FAST cross aggregation, streaming form: one warp walks one whole line.
//
// A pass along a line is out[j] = E[j+hi(j)+1] - E[j-lo(j)] with E the exclusive prefix of the line and lo, hi <= 34
// the pixel's arms (cross_aggregator.cpp:353-389).  Because the arms are bounded, an output only ever needs prefix
// entries within [j-34, j+35]: a warp walks the line ONCE, keeps the running compensated prefix in registers and the
// last 72 entries in a shared-memory ring, and emits output j as soon as the prefix has reached j+35.
//   * lane = disparity: every load / store is a 128-byte row of the [H][W][D] volume; arms, support counts and ring
//     slots are warp-uniform, prepared one chunk (24 positions, one per lane) at a time and read back as broadcast LDS;
//   * no tiles: no halo re-reads, no segment totals, no block barriers -- the ring is lane-private storage, 18.4 KB per
//     warp, which is what lets 12 warps share an SM;
//   * the ring has 72 = 3*24 slots and chunks are 24 positions, so inside a chunk every ring store has a compile-time
//     offset; the interior chunks of a line run a predicate-free path;
//   * a chunk is ordered around the ONE scoreboard ptxas gives all global loads (see seq_chunk): consume what the
//     previous chunk fetched, then issue this chunk's loads, then compute; DRAM -> L2 runs two chunks further ahead with
//     one prefetch instruction per chunk; descriptor and ring reads are software-pipelined over the positions;
//   * per cell: 4 FADD (prefix) + 3 FADD (difference) + 3 FMA-class ops for the exact division + 1 STS.64 + 3 LDS + 1 LDG +
//     1 STG + ~4 integer ops, against ~65 instructions for the tiled kernel in cross_agg_fast.cu.
// Numerics are those of cross_agg_fast.cu: cascaded compensated prefix (s, c) for non-negative input (a power-of-two
// bias keeps the FastTwoSum error term exact), general TwoSum pair otherwise; window sums are exact to ~2^-46 and
// rounded once.  The division by the support count is the correctly rounded quotient (Markstein's residual step on the
// correctly rounded reciprocal), i.e. the same bits as the reference's IEEE `cost / count` (:389).
#include "common.cuh"

// Development switches for the ablation runs quoted in DESIGN.md (all off in the product build):
//   -DSMT_ABL_NOLOAD   volume loads replaced by a runtime constant      -DSMT_ABL_NOSTORE  volume stores predicated off
//   -DSMT_ABL_OCC16    timing only (WRONG results): 32-entry ring and one chunk in flight, so that 16 warps fit per SM
#if defined(SMT_ABL_OCC16)
#define SMT_SEQ_AHEAD 1
#endif
namespace {

#if defined(SMT_ABL_OCC16)
constexpr int SEQ_R = 32;
constexpr int SEQ_RING_SLOTS = 48;
#elif defined(SMT_SEQ_C32)
constexpr int SEQ_RING_SLOTS = 96;
constexpr int SEQ_R = 96;     // three 32-position chunks, 24.6 KB per warp, 8 warps per SM: 0.91 ms per pass
#else
constexpr int SEQ_R = 72;     // ring entries: three 24-position chunks (span needed: 71) = 18.4 KB per warp, so that
constexpr int SEQ_RING_SLOTS = 72;  // 12 warps fit per SM (3 per scheduler): 0.84 ms per pass
#endif
constexpr int SEQ_LOOK = 35;  // output j becomes computable when E[j+35] has been stored
#if defined(SMT_SEQ_C32) || defined(SMT_ABL_OCC16)
constexpr int SEQ_CH = 32;
#else
constexpr int SEQ_CH = 24;    // positions per chunk (one descriptor per lane, lanes 24-31 idle in the descriptor step)
#endif
// Measured on B200 at 1080p x 256, ms per pass (history in DESIGN.md 4): 1.30 with the loads issued at the top of a chunk
// (every chunk then waited for its own loads); 1.18 with the order below; 1.03 with the L2 prefetch; 0.93 with the
// software-pipelined shared-memory reads; 0.82 with 24-position chunks and 12 warps per SM.  Flat at this point: prefetch
// distance 1-6 chunks, lags 3-6, one or two chunks in flight, store cache policy, block size 2-12 warps.
#ifndef SMT_SEQ_AHEAD
#define SMT_SEQ_AHEAD 1
#endif
#ifndef SMT_SEQ_L2PF
#define SMT_SEQ_L2PF 2  // chunks between the register loads and the L2 prefetch running ahead of them (0 = no prefetch)
#endif
constexpr int SEQ_AHEAD = SMT_SEQ_AHEAD;  // chunks of input in flight per warp (register buffers = SEQ_AHEAD + 1)
#ifndef SMT_SEQ_ST
#define SMT_SEQ_ST 0
#endif
__device__ __forceinline__ void seq_store(float* p, float v) {  // cache policy of the volume stores (0 = streaming)
#if SMT_SEQ_ST == 0
    __stcs(p, v);
#elif SMT_SEQ_ST == 1
    *p = v;
#elif SMT_SEQ_ST == 2
    __stcg(p, v);
#else
    __stwt(p, v);
#endif
}
#ifndef SMT_SEQ_PIPE
#define SMT_SEQ_PIPE 1
#endif
#ifndef SMT_SEQ_DLAG
#define SMT_SEQ_DLAG 4
#endif
#ifndef SMT_SEQ_RLAG
#define SMT_SEQ_RLAG 4
#endif
constexpr int SEQ_DLAG = SMT_SEQ_DLAG;   // descriptor read this many positions ahead of its ring reads
constexpr int SEQ_RLAG = SMT_SEQ_RLAG;   // ring reads consumed this many positions after they are issued
constexpr int SEQ_G = 8;      // outputs are emitted in groups of SEQ_G (ring span needed: 70 + SEQ_G <= SEQ_R)

struct pair2 {
    float a, b;
};
template <bool NONNEG>
__device__ __forceinline__ pair2 pre_add(pair2 p, float v) {
    pair2 r;
    if (NONNEG) {  // (s, c): value s + c, FastTwoSum (|s| >= |v| thanks to the bias)
        r.a = __fadd_rn(p.a, v);
        r.b = __fadd_rn(p.b, __fsub_rn(v, __fsub_rn(r.a, p.a)));
    } else {  // (hi, lo): TwoSum + renormalisation
        const float s = __fadd_rn(p.a, v);
        const float bb = __fsub_rn(s, p.a);
        const float e = __fadd_rn(__fsub_rn(p.a, __fsub_rn(s, bb)), __fsub_rn(v, bb));
        const float l = __fadd_rn(p.b, e);
        r.a = __fadd_rn(s, l);
        r.b = __fsub_rn(l, __fsub_rn(r.a, s));
    }
    return r;
}

// correctly rounded a / b for b = a support count (integer 1..4761) given y = RN(1/b)
__device__ __forceinline__ float div_by_count(float a, float b, float y) {
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

// window descriptor of one output position: byte offsets of E[j-lo] and E[j+hi+1] in the lane's ring column, the
// support count as float and its reciprocal
struct __align__(16) Desc {
    unsigned offa, offb;
    float cnt, rcp;
};

// One chunk of 32 positions starting at tb (a multiple of 32).  Loads chunk tb+32 into `vload`, consumes `vuse`
// (chunk tb): 32 prefix steps (E[tb+k] is stored BEFORE v[tb+k] is added, so slots never wrap inside a chunk), and
// after every 8 of them the 8 outputs j = tb+kk-35 that have become computable.  EDGE adds the range predicates.
struct Prefetched {  // raw uchar4 arm word and support count of this lane's output position in the NEXT chunk
    unsigned armw;
    int cnt;
};

__device__ __forceinline__ unsigned f2key_(float f) {  // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f_(unsigned k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }

// WTA: instead of storing the pass output, reduce it over the warp's 32 disparities (first minimum) and store one
// (value, d) pair per position and d-chunk; wta_reduce_kernel finishes the arg-min over the chunks.
template <bool NONNEG, bool HAS_SUP, bool EDGE, bool WTA, int CSTEP = 0>
__device__ __forceinline__ void seq_chunk(pair2& acc, Prefetched& pre, float (&vuse)[SEQ_CH], float (&vload)[SEQ_CH], int tb, int len, bool act,
                                          const float* src_next, float* dst_first, long long step_rt,
                                          long long pbase, long long pstride, int horiz, const uchar4* __restrict__ arms,
                                          const uint16_t* __restrict__ sup, unsigned char* ring_col, Desc* desc, int lane,
                                          float2* __restrict__ wta_first, long long wta_step, int d, int zero, int st_lo, int st_hi) {
    const long long step = CSTEP ? (long long)CSTEP : step_rt;
    const long long step_b = step * 4;  // byte steps on char pointers: one 64-bit add per address instead of index + shift
    // ORDER MATTERS.  ptxas puts every global load of this kernel on ONE scoreboard (SB5 in the SASS), and a scoreboard is a
    // counter: the first instruction that needs ANY loaded value waits until ALL loads issued before it have landed.
    // Issuing this chunk's loads at the top (as a prefetch) therefore made the first prefix step wait for them: a full
    // DRAM round trip per chunk, no overlap at all (28 % long-scoreboard stalls; removing the loads took a pass from
    // 1.30 to 0.82 ms).  So: (1) consume what the previous chunk fetched -- arm word, support count, vuse[0] -- while the only
    // loads in flight are a whole chunk old, (2) only then issue this chunk's loads, with an address that depends on the
    // result of (1) so that the scheduler cannot hoist them, (3) do the rest of the chunk, which needs no loaded value
    // that has not been waited for already.
    // ---- (1a) descriptors of the outputs of this chunk (position tb - 35 + lane) from the arm word / count fetched one chunk ago
    {
        const int j = tb - SEQ_LOOK + lane;
        Desc ds{0u, 0u, 1.f, 1.f};
        if ((SEQ_CH == 32 || lane < SEQ_CH) && (!EDGE || (j >= 0 && j < len))) {
            const unsigned aw = horiz ? (pre.armw & 0xffffu) : (pre.armw >> 16);
            const int lo = (int)(aw & 0xffu), hi = (int)(aw >> 8);
            // arms stop at the image border, so the clamps only act on the halo rows of a row band (bandsplit), whose
            // results are never used
            const int sa = max(j - lo, 0) % SEQ_R, sb = min(j + hi + 1, len) % SEQ_R;
            ds.offa = (unsigned)sa * 256u;
            ds.offb = (unsigned)sb * 256u;
            if (HAS_SUP) {
                ds.cnt = (float)pre.cnt;
                ds.rcp = __frcp_rn(ds.cnt);
            }
        }
        __syncwarp();  // the previous chunk's descriptor reads are done
        if (SEQ_CH == 32 || lane < SEQ_CH) desc[lane] = ds;
        __syncwarp();
    }
    unsigned char* wbase = ring_col + (unsigned)(tb % SEQ_R) * 256u;  // slot of E[tb]
    char* dp = reinterpret_cast<char*>(dst_first);                     // output position tb - 35
    float2* wp = wta_first;
    // ---- (1b) first prefix step
    *reinterpret_cast<float2*>(wbase) = make_float2(acc.a, acc.b);  // E[tb]
    acc = pre_add<NONNEG>(acc, vuse[0]);
    // ---- (2) loads of chunk tb + SEQ_AHEAD*32 and of the next chunk's arm words; `zero` is 0 at run time
    const int dep = __float_as_int(acc.a) & zero;
    {
        const char* p = reinterpret_cast<const char*>(src_next + dep);
#pragma unroll
        for (int k = 0; k < SEQ_CH; ++k) {
            if (EDGE)
                vload[k] = (tb + SEQ_AHEAD * SEQ_CH + k < len && act) ? __ldcs(reinterpret_cast<const float*>(p)) : 0.f;
            else
#if defined(SMT_ABL_NOLOAD)
                vload[k] = __int_as_float(horiz + 0x3f000000);
#else
                vload[k] = __ldcs(reinterpret_cast<const float*>(p));
#endif
            p += step_b;
        }
#if SMT_SEQ_L2PF > 0
        {   // DRAM -> L2 runs further ahead than the scoreboard allows registers to: ONE prefetch instruction per chunk, lane l
            // asks for the 128-byte row of position tb + (AHEAD + L2PF)*32 + l.  The register loads above then hit L2.
            const int q = tb + (SEQ_AHEAD + SMT_SEQ_L2PF) * SEQ_CH + lane;
            if ((SEQ_CH == 32 || lane < SEQ_CH) && q < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(src_next + dep + step * (SMT_SEQ_L2PF * SEQ_CH + lane)));
        }
#endif
        const int jn = tb - SEQ_LOOK + lane + SEQ_CH;
        pre.armw = 0u;
        pre.cnt = 1;
        if ((SEQ_CH == 32 || lane < SEQ_CH) && jn >= 0 && jn < len) {
            const long long pix = pbase + pstride * jn + dep;
            pre.armw = __ldg(reinterpret_cast<const unsigned*>(arms) + pix);  // raw {left,right,top,bottom} bytes
            if (HAS_SUP) pre.cnt = (int)__ldg(sup + pix);
        }
    }
    // ---- (3) the other prefix steps and the outputs, software-pipelined over the 32 positions (everything is unrolled, the
    // arrays are registers): the descriptor of output k is read SEQ_DLAG steps before its two ring reads, which are issued
    // right after the prefix step that completes them (E[tb+k] is the newest entry output k can need) and consumed
    // SEQ_RLAG steps later.  Left alone the compiler serialises descriptor read -> ring reads -> arithmetic per output and
    // every output pays two shared-memory round trips (ncu: 32 % short-scoreboard stalls).
#if SMT_SEQ_PIPE
    {
        Desc dsc[SEQ_CH];
        float2 ea[SEQ_CH], eb[SEQ_CH];
#pragma unroll
        for (int u = 0; u < SEQ_DLAG; ++u) dsc[u] = desc[u];  // broadcast reads
#pragma unroll
        for (int k = 0; k < SEQ_CH + SEQ_RLAG; ++k) {
            if (k < SEQ_CH) {
                if (k > 0) {
                    *reinterpret_cast<float2*>(wbase + k * 256) = make_float2(acc.a, acc.b);  // E[tb+k]
                    acc = pre_add<NONNEG>(acc, vuse[k]);
                }
                ea[k] = *reinterpret_cast<const float2*>(ring_col + dsc[k].offa);
                eb[k] = *reinterpret_cast<const float2*>(ring_col + dsc[k].offb);
                if (k + SEQ_DLAG < SEQ_CH) dsc[k + SEQ_DLAG] = desc[k + SEQ_DLAG];
            }
            if (k >= SEQ_RLAG) {
                const int kk = k - SEQ_RLAG;
                const bool inside = !EDGE || (tb - SEQ_LOOK + kk >= st_lo && tb - SEQ_LOOK + kk < st_hi);  // warp-uniform
                float r = __fadd_rn(__fsub_rn(eb[kk].x, ea[kk].x), __fsub_rn(eb[kk].y, ea[kk].y));
                if (HAS_SUP) r = div_by_count(r, dsc[kk].cnt, dsc[kk].rcp);
                if (WTA) {
                    if (inside) {
                        const unsigned key = act ? f2key_(r) : 0xffffffffu;
                        const unsigned g = __reduce_min_sync(0xffffffffu, key);
                        const unsigned who = __ballot_sync(0xffffffffu, key == g);
                        if (lane == 0) *wp = make_float2(key2f_(g), (float)(d + __ffs(who) - 1));  // d = this warp's first disparity
                    }
                } else if (inside && (!EDGE || act)) {
#if defined(SMT_ABL_NOSTORE)
                    if (r == 123456.f)
#endif
                    seq_store(reinterpret_cast<float*>(dp), r);
                }
                dp += step_b;
                wp += wta_step;
            }
        }
    }
#else
#pragma unroll
    for (int k = 0; k < SEQ_CH; ++k) {
        if (k > 0) {
            *reinterpret_cast<float2*>(wbase + k * 256) = make_float2(acc.a, acc.b);  // E[tb+k]
            acc = pre_add<NONNEG>(acc, vuse[k]);
        }
        if ((k & (SEQ_G - 1)) == SEQ_G - 1) {
#pragma unroll
            for (int u = 0; u < SEQ_G; ++u) {
                const int kk = k - (SEQ_G - 1) + u;
                const Desc ds = desc[kk];  // broadcast
                if (!EDGE || (tb - SEQ_LOOK + kk >= st_lo && tb - SEQ_LOOK + kk < st_hi)) {
                    const float2 ea = *reinterpret_cast<const float2*>(ring_col + ds.offa);
                    const float2 eb = *reinterpret_cast<const float2*>(ring_col + ds.offb);
                    float r = __fadd_rn(__fsub_rn(eb.x, ea.x), __fsub_rn(eb.y, ea.y));
                    if (HAS_SUP) r = div_by_count(r, ds.cnt, ds.rcp);
                    if (WTA) {
                        const unsigned key = act ? f2key_(r) : 0xffffffffu;
                        const unsigned g = __reduce_min_sync(0xffffffffu, key);
                        const unsigned who = __ballot_sync(0xffffffffu, key == g);
                        if (lane == 0) *wp = make_float2(key2f_(g), (float)(d + __ffs(who) - 1));  // d = this warp's first disparity
                    } else if (!EDGE || act) {
#if defined(SMT_ABL_NOSTORE)
                        if (r == 123456.f)
#endif
                        __stcs(reinterpret_cast<float*>(dp), r);
                    }
                }
                dp += step_b;
                wp += wta_step;
            }
        }
    }
#endif
}

#if defined(SMT_SEQ_C32) || defined(SMT_ABL_OCC16)
constexpr int SEQ_WARPS = 8;
#elif defined(SMT_SEQ_WARPS)
constexpr int SEQ_WARPS = SMT_SEQ_WARPS;
#else
constexpr int SEQ_WARPS = 4;   // consecutive (line, d-chunk) tasks; three such blocks share an SM (12 warps), small blocks
                               // keep small images spread over all SMs (12-warp blocks: equal at 1080p, 20 % slower at 450x375)
#endif
constexpr int SEQ_CTAS_PER_SM = 12 / SEQ_WARPS > 0 ? 12 / SEQ_WARPS : 1;
constexpr int SEQ_SMEM_PER_WARP = SEQ_RING_SLOTS * 256;  // the descriptors live in a static array

template <bool NONNEG, bool HAS_SUP, bool WTA, int CSTEP = 0>
__global__ void __launch_bounds__(SEQ_WARPS * 32, SEQ_CTAS_PER_SM) agg_seq_kernel(const float* in, float* out,  // may alias: a redo of a fused pair runs its first pass in place
                                                    
                                                     const uchar4* __restrict__ arms, const uint16_t* __restrict__ sup, int H, int W,
                                                     int D, int horiz, float bias, float2* __restrict__ wta_partial, int zero, int tmaps,
                                                     const int* __restrict__ redo, int nseg, int st_lo, int st_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* s_ring = smem_raw + (size_t)warp * SEQ_SMEM_PER_WARP;  // [slot][lane] float2
    // a separate object, so that the compiler may move descriptor loads across ring stores
    __shared__ Desc s_desc_all[SEQ_WARPS][SEQ_CH];
    Desc* s_desc = s_desc_all[warp];
    const int nchunk = (D + 31) >> 5;
    const int len = horiz ? W : H;
    const int st_lo0 = st_lo, st_hi0 = st_hi;
    // redo != nullptr: this launch is queued behind a fixed-point sweep (cross_agg_fx.cu) and only redoes the (line, chunk)
    // tasks that sweep put on its list -- or every task once the sweep escalated.  The grid has one warp per task of the
    // sweep; listed (line, segment) items are spread over its warps with a stride (there can be more items than warps),
    // warps beyond the list leave at once, and any other launch leaves the loop after its first round.
    // A listed line is cut into nseg segments walked by different warps: a handful of warps each walking a whole line
    // would leave the launch bound by the latency of one walk (~50 us at 1920 positions).  A segment starts its prefix
    // 48 positions early (outputs only read prefix entries up to 34 positions back) and stores its own positions only.
    // Not for in-place launches (in == out): a segment's lead-in would read what its neighbour has already overwritten.
    const bool listed = redo && *reinterpret_cast<const volatile int*>(redo + FX_ST_ESCALATED) == 0;
    const long long nlisted = listed ? (long long)*reinterpret_cast<const volatile int*>(redo + FX_ST_COUNT) * nseg : 0;
    for (long long v = (long long)blockIdx.x * SEQ_WARPS + warp;; v += (long long)gridDim.x * SEQ_WARPS) {
    long long wg = v;
    int seg_lo = 0, seg_hi = 0x7fffffff;
    if (listed) {
        if (v >= nlisted) break;
        const long long idx = v / nseg;
        wg = redo[FX_ST_LIST + idx];
        if (nseg > 1) {
            const int seg_len = (len + nseg - 1) / nseg;
            seg_lo = (int)(v - idx * nseg) * seg_len;
            seg_hi = seg_lo + seg_len;
        }
    }
    const int line = (int)(wg / nchunk), chunk = (int)(wg - (long long)line * nchunk);
    if (line >= (horiz ? H : W)) break;  // whole warp; warps never synchronise with each other
    if (seg_lo >= len) continue;
    st_lo = max(st_lo0, seg_lo);
    st_hi = min(st_hi0, seg_hi);
    const int o_first = max(0, (seg_lo - 48) / SEQ_CH);                                      // first chunk walked
    const int o_last = seg_hi >= len ? 0x7fffffff : (seg_hi + SEQ_LOOK + SEQ_CH - 1) / SEQ_CH;  // one past the last
    const int d = chunk * 32 + lane;
    const bool act = d < D;
    const bool full = chunk * 32 + 32 <= D;  // every lane active: the interior path needs no lane predicate
    const long long pstride = horiz ? 1 : (long long)W;
    const long long pbase = horiz ? (long long)line * W : (long long)line;
    const long long step = CSTEP ? (long long)CSTEP : pstride * D;
    const float* src = in + pbase * D + (act ? d : 0);
    float* dst = out + pbase * D + (act ? d : 0);
    unsigned char* ring_col = s_ring + lane * 8;
    // index of position j of this line in the arm / support-count maps: [H][W], or [W][H] copies for vertical walks (tmaps)
    // so that the 32 lanes of a descriptor fetch read consecutive entries instead of 32 different rows
    const long long mstride = (!horiz && tmaps) ? 1 : pstride;
    const long long mbase = (!horiz && tmaps) ? (long long)line * H : pbase;

    pair2 acc{NONNEG ? bias : 0.f, 0.f};
    // SEQ_AHEAD + 1 register buffers rotate: while chunk o is consumed, the next SEQ_AHEAD chunks are in flight
    float va[SEQ_CH], vb[SEQ_CH], vc[SEQ_AHEAD == 2 ? SEQ_CH : 1];
#pragma unroll
    for (int k = 0; k < SEQ_CH; ++k) {
        const int p = o_first * SEQ_CH + k;
        va[k] = (p < len && act) ? __ldcs(src + step * p) : 0.f;
        if (SEQ_AHEAD == 2) vb[k] = (SEQ_CH + p < len && act) ? __ldcs(src + step * (SEQ_CH + p)) : 0.f;
    }

    Prefetched pre{0u, 1};  // chunk 0 outputs positions -35..-4: nothing to fetch; its call fetches chunk 1's
    // chunk o covers prefix positions [32*o, 32*o+32) and outputs [32*o-35, 32*o-3); outputs end at len-1
    const int nouter = min((len + SEQ_LOOK + SEQ_CH - 1) / SEQ_CH, o_last);
#define SMT_SEQ_STEP(USE, LOAD)                                                                                               \
    if (o < nouter) {                                                                                                         \
        const int tb = o * SEQ_CH;                                                                                            \
        const float* src_next = src + step * (tb + SEQ_AHEAD * SEQ_CH);                                                       \
        float* dst_first = dst + step * (tb - SEQ_LOOK);                                                                      \
        float2* wta_first = WTA ? wta_partial + (pbase + pstride * (tb - SEQ_LOOK)) * nchunk + chunk : nullptr;               \
        const long long wta_step = pstride * nchunk;                                                                          \
        /* interior: the chunk being loaded is inside the line, all outputs of this chunk are, all lanes active */           \
        if (full && tb - SEQ_LOOK >= st_lo && tb - SEQ_LOOK + SEQ_CH <= st_hi && tb + (SEQ_AHEAD + 1) * SEQ_CH <= len)        \
            seq_chunk<NONNEG, HAS_SUP, false, WTA, CSTEP>(acc, pre, USE, LOAD, tb, len, act, src_next, dst_first, step, mbase,       \
                                                   mstride, horiz, arms, sup, ring_col, s_desc, lane, wta_first,               \
                                                   wta_step, chunk * 32, zero, st_lo, st_hi);                                 \
        else                                                                                                                  \
            seq_chunk<NONNEG, HAS_SUP, true, WTA, CSTEP>(acc, pre, USE, LOAD, tb, len, act, src_next, dst_first, step, mbase,        \
                                                  mstride, horiz, arms, sup, ring_col, s_desc, lane, wta_first,                \
                                                  wta_step, chunk * 32, zero, st_lo, st_hi);                                  \
        ++o;                                                                                                                  \
    }
    for (int o = o_first; o < nouter;) {
#if SMT_SEQ_AHEAD == 2
        SMT_SEQ_STEP(va, vc)
        SMT_SEQ_STEP(vb, va)
        SMT_SEQ_STEP(vc, vb)
#else
        SMT_SEQ_STEP(va, vb)
        SMT_SEQ_STEP(vb, va)
#endif
    }
    }
#undef SMT_SEQ_STEP
}

template <bool NONNEG, bool HAS_SUP, bool WTA>
static smt_status_t launch_seq(smt_ctx* ctx, const float* in, float* out, const uchar4* arms, const uint16_t* sup, int H, int W, int D,
                               int horiz, float bias, float2* wta_partial = nullptr, int tmaps = 0, const int* redo = nullptr, int st_lo = 0,
                               int st_hi = 0) {
    // redo launches that do not run in place cut the listed lines into segments (see the kernel)
#ifndef SMT_REDO_SEGS
#define SMT_REDO_SEGS 8
#endif
    const int nseg = (redo && in != out) ? max(1, min(SMT_REDO_SEGS, (horiz ? W : H) / 192)) : 1;
    if (st_hi <= 0) {
        st_lo = 0;
        st_hi = horiz ? W : H;
    }
    constexpr size_t smem = (size_t)SEQ_WARPS * SEQ_SMEM_PER_WARP;
    static bool attr_set_dev[64] = {};  // function attributes are per device
    bool& attr_set = attr_set_dev[ctx->device & 63];
    if (!attr_set) {
        SMT_CUDA(ctx, cudaFuncSetAttribute(agg_seq_kernel<NONNEG, HAS_SUP, WTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // three blocks per SM need 225 of the 228 KB: ask for the largest shared-memory carve-out explicitly
        SMT_CUDA(ctx, cudaFuncSetAttribute(agg_seq_kernel<NONNEG, HAS_SUP, WTA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           (int)cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    const long long nwarps = (long long)(horiz ? H : W) * ((D + 31) / 32);
    // One warp per task, also for redo launches: warps beyond the list leave at once.  (A small fixed grid whose warps stride
    // over the list measured the same on listed sweeps and 11 % slower on escalated whole sweeps at 608 x 7680 x 512, where
    // all warps then walk in lockstep; the task loop in the kernel is what is left of it.)
    // horizontal walk of a 256-disparity volume (the benchmark shape): the position step is a compile-time constant, so
    // the 32 loads and 32 stores of a chunk use immediate offsets from one base register (-27 % instructions, +1.3 %)
    if constexpr (NONNEG && !WTA) if (horiz && D == 256) {
        static bool attr2_dev[64] = {};
        bool& attr2 = attr2_dev[ctx->device & 63];
        if (!attr2) {
            SMT_CUDA(ctx, cudaFuncSetAttribute(agg_seq_kernel<NONNEG, HAS_SUP, WTA, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SMT_CUDA(ctx, cudaFuncSetAttribute(agg_seq_kernel<NONNEG, HAS_SUP, WTA, 256>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                               (int)cudaSharedmemCarveoutMaxShared));
            attr2 = true;
        }
        SMT_LAUNCH(ctx, (agg_seq_kernel<NONNEG, HAS_SUP, WTA, 256>), dim3(cdiv(nwarps, SEQ_WARPS)), SEQ_WARPS * 32, smem, in, out, arms, sup, H, W, D,
                   horiz, bias, wta_partial, 0, tmaps, redo, nseg, st_lo, st_hi);
        return SMT_OK;
    }
    SMT_LAUNCH(ctx, (agg_seq_kernel<NONNEG, HAS_SUP, WTA>), dim3(cdiv(nwarps, SEQ_WARPS)), SEQ_WARPS * 32, smem, in, out, arms, sup, H, W, D, horiz,
               bias, wta_partial, 0, tmaps, redo, nseg, st_lo, st_hi);
    return SMT_OK;
}

}  // namespace

smt_status_t k_agg_seq_pass(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                            int horiz, bool nonneg, float bias, int tmaps, const int* redo, int st_lo, int st_hi) {
    const uchar4* arms = reinterpret_cast<const uchar4*>(arms8);
#ifdef SMT_DEV_NO_REDO  // timing experiment only: what the redo launches behind the fixed-point sweeps cost (results are wrong)
    if (redo) return SMT_OK;
#endif
    if (nonneg)
        return sup ? launch_seq<true, true, false>(ctx, in, out, arms, sup, H, W, D, horiz, bias, nullptr, tmaps, redo, st_lo, st_hi)
                   : launch_seq<true, false, false>(ctx, in, out, arms, sup, H, W, D, horiz, bias, nullptr, tmaps, redo, st_lo, st_hi);
    return sup ? launch_seq<false, true, false>(ctx, in, out, arms, sup, H, W, D, horiz, bias, nullptr, tmaps, redo, st_lo, st_hi)
               : launch_seq<false, false, false>(ctx, in, out, arms, sup, H, W, D, horiz, bias, nullptr, tmaps, redo, st_lo, st_hi);
}

// [H][W] -> [W][H] copies of the arm words and of one support-count map for the vertical walks
__global__ void __launch_bounds__(256) transpose_maps_kernel(const uchar4* __restrict__ arms, const uint16_t* __restrict__ sup, int H, int W,
                                                             uchar4* __restrict__ armsT, uint16_t* __restrict__ supT) {
    __shared__ unsigned ta[32][33];
    __shared__ uint16_t ts[32][34];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int y = y0 + r, x = x0 + tx;
        if (y < H && x < W) {
            ta[r][tx] = reinterpret_cast<const unsigned*>(arms)[(size_t)y * W + x];
            ts[r][tx] = sup ? sup[(size_t)y * W + x] : (uint16_t)0;
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + r, y = y0 + tx;
        if (y < H && x < W) {
            reinterpret_cast<unsigned*>(armsT)[(size_t)x * H + y] = ta[tx][r];
            if (supT) supT[(size_t)x * H + y] = ts[tx][r];
        }
    }
}
smt_status_t k_transpose_maps(smt_ctx* ctx, const uint8_t* arms8, const uint16_t* sup, int H, int W, uint8_t* armsT, uint16_t* supT) {
    dim3 grid(cdiv(W, 32), cdiv(H, 32));
    SMT_LAUNCH(ctx, transpose_maps_kernel, grid, 256, 0, reinterpret_cast<const uchar4*>(arms8), sup, H, W,
               reinterpret_cast<uchar4*>(armsT), supT);
    return SMT_OK;
}

// one sweep's redo state: the list is emptied; a new aggregation also clears the escalation and the running total
smt_status_t k_fx_state_reset(smt_ctx* ctx, int* state, bool new_aggregation, size_t ntasks) {
    const int limit = (int)(ntasks / 4);
    if (new_aggregation) {
        const int init[4] = {0, 0, limit, 0};
        for (int k = 0; k < 4; ++k) ctx->h_pinned[32 + k] = init[k];
        SMT_CUDA(ctx, cudaMemcpyAsync(state, ctx->h_pinned + 32, 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        SMT_CUDA(ctx, cudaMemsetAsync(state + FX_ST_COUNT, 0, sizeof(int), ctx->stream));
    }
    return SMT_OK;
}

smt_status_t k_cross_sweep(smt_ctx* ctx, int iters, int sweep, const float* vol_in, float* vol_out, const uint8_t* arms,
                           const uint16_t* sup_hfirst, const uint16_t* sup_vfirst, int rows, int cols, int D, float vmax,
                           smt_sum_mode_t mode, bool first_of_aggregation, int st_lo, int st_hi, size_t ntasks_whole) {
    (void)iters;
    StageTimer t(ctx, ST_AGGREGATE);
    // direction and divisor of sweep s (cross_aggregator.cpp:102-117): horizontal first on even iterations, second pass divides
    const bool hfirst = (sweep / 2) % 2 == 0, second = sweep % 2 == 1;
    const int horiz = (hfirst != second) ? 1 : 0;
    const uint16_t* sup = second ? (hfirst ? sup_hfirst : sup_vfirst) : nullptr;
    if (mode == SMT_SUM_REFERENCE_ORDER) return k_agg_exact_pass(ctx, vol_in, vol_out, arms, sup, rows, cols, D, horiz, st_lo, st_hi);
    const bool nonneg = vmax > 0.f && vmax < 1e30f;
    float b = 1.f;
    while (nonneg && b < vmax) b *= 2.f;
    float inv_q = 0.f, q = 0.f;
    const bool fx = nonneg && fx_quantum(vmax, &inv_q, &q);
    size_t ntasks = (size_t)(rows > cols ? rows : cols) * ((D + 31) / 32);
    if (ntasks_whole > ntasks) ntasks = ntasks_whole;
    int* state = fx ? smt_redo_state(ctx, ntasks) : nullptr;
    if (fx && !state) return smt_fail(ctx, SMT_ERR_OOM, "cross sweep: redo list allocation failed");
    if (fx) {
        SMT_TRY(k_fx_state_reset(ctx, state, first_of_aggregation, ntasks));
        SMT_TRY(k_agg_fx_pass(ctx, vol_in, vol_out, arms, sup, rows, cols, D, horiz, 0, inv_q, q, state, st_lo, st_hi));
    }
    return k_agg_seq_pass(ctx, vol_in, vol_out, arms, sup, rows, cols, D, horiz, nonneg, second ? b * 128.f : b, 0, state, st_lo, st_hi);
}

// ---------------------------------------------------------------------------------------------- per-sweep C ABI (row bands)
extern "C" {

int32_t smt_cross_sweep_count(int32_t iters) { return iters > 0 ? 2 * iters : 0; }

// sweep s of `iters` iterations: direction and whether it divides by a support count (cross_aggregator.cpp:102-117)
smt_status_t smt_cross_sweep_info(int32_t iters, int32_t sweep, int32_t* horizontal, int32_t* divides) {
    if (sweep < 0 || sweep >= 2 * iters || !horizontal || !divides) return SMT_ERR_INVALID_ARG;
    const bool hfirst = (sweep / 2) % 2 == 0;
    const bool second = sweep % 2 == 1;
    *horizontal = (hfirst != second) ? 1 : 0;
    *divides = second ? 1 : 0;
    return SMT_OK;
}

smt_status_t smt_aggregate_cross_sweep(smt_ctx* ctx, int32_t iters, int32_t sweep, const float* vol_in, float* vol_out,
                                       const uint8_t* arms, const uint16_t* sup_hfirst, const uint16_t* sup_vfirst, int32_t rows,
                                       int32_t cols, int32_t D, float vmax, smt_sum_mode_t mode, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, space == SMT_DEVICE, SMT_ERR_UNSUPPORTED, "smt_aggregate_cross_sweep: device memspace only");
    SMT_REQUIRE(ctx, vol_in && vol_out && vol_in != vol_out && arms && sup_hfirst && sup_vfirst && rows > 0 && cols > 0 && D > 0,
                SMT_ERR_INVALID_ARG, "smt_aggregate_cross_sweep: bad argument");
    SMT_REQUIRE(ctx, sweep >= 0 && sweep < 2 * iters, SMT_ERR_INVALID_ARG, "smt_aggregate_cross_sweep: sweep outside 0..2*iters-1");
    SMT_REQUIRE(ctx, mode == SMT_SUM_REFERENCE_ORDER || mode == SMT_SUM_FAST, SMT_ERR_INVALID_ARG,
                "smt_aggregate_cross_sweep: unknown sum mode");
    DeviceGuard dev_guard(ctx->device);
    if (sweep == 0 && mode == SMT_SUM_FAST) {
        // the streaming kernels keep 72 prefix entries per line: arms beyond 34 would wrap onto stale entries
        int longest = 0;
        SMT_TRY(k_arms_max(ctx, arms, (size_t)rows * cols, &longest));
        if (longest > 34)
            return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "smt_aggregate_cross_sweep: SMT_SUM_FAST supports arms <= 34, the maps hold %d", longest);
    }
    return k_cross_sweep(ctx, iters, sweep, vol_in, vol_out, arms, sup_hfirst, sup_vfirst, rows, cols, D, vmax, mode, sweep == 0);
}

smt_status_t smt_aggregate_fast_status(smt_ctx* ctx, int32_t* used_float_pairs) {
    if (!ctx || !used_float_pairs) return SMT_ERR_INVALID_ARG;
    DeviceGuard dev_guard(ctx->device);
    *used_float_pairs = 0;
    if (!ctx->d_redo) return SMT_OK;  // no fast aggregation has run on this context
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_redo, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *used_float_pairs = ctx->h_pinned[FX_ST_ESCALATED] ? -1 : ctx->h_pinned[FX_ST_TOTAL];
    return SMT_OK;
}

}  // extern "C"
