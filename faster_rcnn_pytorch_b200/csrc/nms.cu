// N1: greedy NMS over score-sorted boxes (torchvision.ops.nms semantics, models/model.py:53-55).
//
// Keep-list algorithm on one thread-block cluster per image (no n x n mask in HBM):
//   the cluster walks the sorted candidates in chunks of 256; each CTA owns a round-robin slice
//   of the boxes kept so far (shared memory) and tests the chunk against its slice; the
//   per-CTA suppression words are exchanged through distributed shared memory (one
//   cluster barrier per chunk); every CTA then resolves the chunk's survivors among
//   themselves (warp-ballot predecessor masks + a parallel fix-point) and appends the newly
//   kept boxes to its slice.  The walk stops at max_keep, so with the RPN's 12000 -> 2000
//   only the first ~5-6k candidates are ever touched and each is tested against kept boxes
//   only (~5 M pair tests instead of the 72 M of the full upper triangle).
//
// Phase 1 is register tiled: a thread holds 4 candidates and streams kept boxes from shared
// memory (one broadcast LDS.128 feeds 128 pair tests), because the un-tiled version was bound
// by shared-memory wavefronts, not by the FP32/ALU pipes.
//
// Bit-exactness vs the CPU kernel: IoU = inter / ((area_i + area_j) - inter) in fp32 with IEEE
// division and no FMA contraction; suppress when (double)iou > thr  <=>  iou >= up, where up is
// the smallest fp32 strictly above thr.  Almost every pair is rejected without the division by
// a conservative 10-instruction screen:  inter/union < up  <=  inter < c2*(area_a + area_b) with
// c2 = up/(1+up)*(1 - 2^-19)  (the 2^-19 margin covers every fp32 rounding in the screen and in
// the exact formula, see suppress_screen<kUnit>()); the ~0.3 % of pairs that pass the screen take the exact
// division.  Degenerate / tiny boxes carry a NaN screening area, which always fails the screen.
#include <cooperative_groups.h>
#include <math.h>

#include "frr_common.cuh"

namespace cg = cooperative_groups;

namespace frr {

constexpr int kChunk = 256;  // candidates per chunk
constexpr int kChunkWords = kChunk / 32;
constexpr int kMaxCluster = 16;
constexpr int kTile = 4;                // candidates per thread in phase 1
constexpr int kGroup = kChunk / kTile;  // threads that together cover one chunk (64)

struct NmsThr {
    float up;  // smallest fp32 with (double)up > thr
    float c2;  // up/(1+up) * (1 - 2^-19): screening constant
    float alo, ahi;  // a box of area A can only be suppressed by boxes with area in [alo * A, ahi * A] (IoU <= min/max area)
    float fx;        // ... and only by boxes whose x-centre is within fx * (its width) of its own
    int fast;  // screening usable (1e-6 <= thr, finite)
};

constexpr int kStrips = 88;         // area classes of the sorted kept slice (4 per octave, 2^-22 .. 1)
constexpr int kXBins = 8;           // x-centre bins inside an area class
constexpr int kKeys = kStrips * kXBins;  // bucket keys; key kKeys = boxes that must always be tested
constexpr int kKeyPer = ((kKeys + 2 + 31) / 32 + 3) & ~3;  // keys scanned per lane (a multiple of 4: uint4 accesses)
constexpr int kKeyCap = 32 * kKeyPer;                       // padded length of the per-key arrays

__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// c2-scaled area used by the screen: NaN for boxes that are not well formed or tiny (forces the exact path)
__device__ __forceinline__ float screen_area(const float4& b, float c2) {
    const float a = box_area(b);
    const bool ok = (b.z >= b.x) && (b.w >= b.y) && (a <= 3.0e38f) && (a >= 1.0e-30f);
    return ok ? __fmul_rn(c2, a) : __int_as_float(0x7fc00000);
}

// The exact torchvision CPU decision for one pair (a = earlier box).  Rare path.
__device__ __noinline__ bool suppress_exact(float4 a, float4 b, float up) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(box_area(a), box_area(b)), inter);
    const float ovr = __fdiv_rn(inter, uni);
    return ovr >= up;  // false for NaN, as (double)NaN > thr
}

// Screen: returns false only when the pair is certainly NOT suppressed.
//   exact:  ovr = RN(I / U),  U = RN(RN(Aa+Ab) - I) = (Aa+Ab-I)(1+e), |e| <= 2^-22 (I <= (Aa+Ab)/2)
//   ovr < up  <=  I/U < up(1-2^-23)  <=  I < u'(Aa+Ab-I), u' = up(1-2^-21)  <=>  I < u'/(1+u') (Aa+Ab)
//   screen:  T = RN(RN(c2 Aa) + RN(c2 Ab)) <= c2 (Aa+Ab)(1+2^-22), and c2 (1+2^-22) < u'/(1+u').
// Only one of w/h is clamped: if w < 0 then I <= 0 < T (the true intersection is 0: not suppressed).
// sa/sb are the c2-scaled areas (NaN if degenerate -> T is NaN -> the screen reports "maybe").
// kUnit: all coordinates lie in [0,1] (RPN / detection boxes are clamped there), so |h| <= 1 and the clamp of h at 0
// is the free .sat modifier of the subtraction (FMA pipe) instead of an FMNMX on the half-rate ALU pipe.
template <bool kUnit>
__device__ __forceinline__ bool suppress_screen(const float4& a, float sa, const float4& b, float sb) {
    const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float hd = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float h = kUnit ? __saturatef(hd) : fmaxf(0.f, hd);
    return !(__fmul_rn(w, h) < __fadd_rn(sa, sb));
}

struct NmsSmem {
    unsigned int supp[2][kMaxCluster][kChunkWords];  // [parity][src rank][word]: suppressed-by-slice words
    unsigned int acc[2][kChunkWords];                // [parity][word]: this CTA's words (OR over its parts)
    unsigned int warp_cnt[kChunkWords];
    unsigned int kept_w[2][kChunkWords];  // fix-point state, double buffered
    unsigned int removed_w[2][kChunkWords];
    unsigned int pred[kChunk][kChunkWords];  // pred[i][q]: bit j set if survivor 32q+j (< i) suppresses survivor i
    float4 cbox[kChunk];                     // the chunk's candidates
    float carea[kChunk];
    float4 sbox[kChunk];  // survivor boxes (compacted)
    float sarea[kChunk];
    short ssrc[kChunk];  // survivor -> position inside the chunk
    // sorted phase 1 (kUnit): kept slice bucketed by area class, chunk candidates ordered by area class
    alignas(16) unsigned int shist[kKeyCap];
    alignas(16) unsigned int scursor[kKeyCap];
    alignas(16) unsigned int chist[kKeyCap];
    alignas(16) unsigned int ccursor[kKeyCap];
    alignas(16) unsigned short sstart[kKeyCap];  // sstart[s] = first sorted position of key s; [kKeys+1] = total
    unsigned short cord[kChunk];         // chunk positions ordered by area class
};

// Bucket key of a box = (area class, x bin).
// Area class: the top bits of the fp32 area (exponent + 2 mantissa bits = 4 classes per octave), an exactly monotone
// integer function of the area.  IoU <= min(area) / max(area) (in fp32 as well: w <= both widths and RN is monotone, so
// inter <= both areas), hence only kept boxes whose area lies within [thr, 1/thr] of the candidate's can suppress it: for
// RPN proposals (three anchor scales, a factor 4 apart in area) that alone removes 3/4 of the pairs.
// x bin: kXBins equal strips of the x-centre; IoU >= thr also needs |cx_K - cx_c| <= fx * w_c, which is sharp exactly
// where the area cut is not -- the many small boxes of the most populated classes.
// Boxes without a usable screening area (degenerate / malformed, NaN sa) get the extra key kKeys and are tested
// against everything.
__device__ __forceinline__ int strip_of_area(float a) {
    return min(kStrips - 1, max(0, (__float_as_int(a) >> 21) - ((127 - 22) << 2)));
}
__device__ __forceinline__ int xbin_of(float cx) { return min(kXBins - 1, max(0, (int)(cx * (float)kXBins))); }
__device__ __forceinline__ int strip_of(const float4& b, float sa) {
    return (sa != sa) ? kKeys : strip_of_area(box_area(b)) * kXBins + xbin_of(0.5f * (b.x + b.z));
}

// barrier over the first `n` threads of the CTA with an OR reduction of `pred`
__device__ __forceinline__ bool bar_or(int id, int n, bool pred) {
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"((int)pred), "r"(id), "r"(n)
        : "memory");
    return r != 0;
}

enum { DBG_CHUNKS = 0, DBG_LOAD, DBG_P1, DBG_SYNC, DBG_P2, DBG_P3, DBG_P4, DBG_P5, DBG_SURV, DBG_ITERS, DBG_N };

template <int kThreads, bool kFast, bool kUnit, bool kSort>
__global__ void __launch_bounds__(kThreads)
    nms_keeplist_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ counts, int n, int max_keep,
                        int slice_cap, NmsThr thr, int32_t* __restrict__ keep, int32_t* __restrict__ keep_count,
                        float4* __restrict__ out_boxes, long long* __restrict__ dbg,
                        const int32_t* __restrict__ gather_idx, int src_n) {
    static_assert(kThreads >= kChunk && kThreads % kGroup == 0, "bad thread count");
    constexpr int kWarps = kThreads / 32;
    constexpr int kParts = kThreads / kGroup;  // interleaved parts of the kept slice
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int img = blockIdx.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = tid / kGroup;  // which interleaved part of the kept slice this thread scans
    const int u = tid % kGroup;     // candidates u, u+64, u+128, u+192 of the chunk

    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem* sm = reinterpret_cast<NmsSmem*>(smem_raw);
    float4* kbox = reinterpret_cast<float4*>(smem_raw + ((sizeof(NmsSmem) + 15) & ~(size_t)15));  // [slice_cap]
    float* karea = reinterpret_cast<float*>(kbox + slice_cap);                                      // [slice_cap]
    // kUnit: second buffer for the per-chunk re-bucketing of the slice (layout: boxA | boxB | areaA | areaB)
    constexpr bool kSorted = kFast && kUnit && kSort;
    float4* kbox_alt = kbox + slice_cap;
    if (kSorted) karea = reinterpret_cast<float*>(kbox + 2 * slice_cap);
    float* karea_alt = karea + slice_cap;
    int ns_sorted = 0;  // slice entries already bucketed in the current buffer

    const int cnt = counts ? min(counts[img], n) : n;
    // candidate i of the image: boxes[img][i], or boxes[img][gather_idx[img][i]] when the sorted order is given as
    // indices into an unsorted [B, src_n, 4] array (the top-k kernel then never writes the 12 000 gathered boxes, and
    // only the ~5 000 candidates NMS really visits are ever gathered)
    const float4* ib = boxes + (size_t)img * (gather_idx ? src_n : n);
    const int32_t* gi = gather_idx ? gather_idx + (size_t)img * n : nullptr;
    auto cand = [&](int i) -> float4 { return gi ? ib[gi[i]] : ib[i]; };
    int32_t* ikeep = keep + (size_t)img * max_keep;
    float4* iout = out_boxes ? out_boxes + (size_t)img * max_keep : nullptr;
    const bool prof = (dbg != nullptr) && blockIdx.x == 0 && tid == 0;
    long long t0 = 0;
#define FRR_TICK(slot)                  \
    if (prof) {                         \
        const long long t1 = clock64(); \
        dbg[slot] += t1 - t0;           \
        t0 = t1;                        \
    }

    int nk = 0;  // kept so far (identical in every CTA of the cluster)
    int par = 0;
    float4 nbx = make_float4(0.f, 0.f, 0.f, 0.f);  // prefetched candidate of the next chunk (first kChunk threads)
    if (tid < kChunk && tid < cnt) nbx = cand(tid);
    if (tid < 2 * kChunkWords) (&sm->acc[0][0])[tid] = 0u;
    if (kSorted)
        for (int e = tid; e < kKeys + 2; e += kThreads) sm->sstart[e] = 0;
    for (int base = 0; base < cnt && nk < max_keep; base += kChunk, par ^= 1) {
        if (prof) { t0 = clock64(); dbg[DBG_CHUNKS] += 1; }
        // ---- phase 0: stage the chunk's candidates in shared memory, prefetch the next chunk ----------
        if (tid < kChunk) {
            sm->cbox[tid] = nbx;
            sm->carea[tid] = screen_area(nbx, thr.c2);
            const int nx = base + kChunk + tid;
            if (nx < cnt) nbx = cand(nx);
        }
        __syncthreads();
        FRR_TICK(DBG_LOAD);

        const int ns = (nk - rank + S - 1) / S;  // kept ordinals o with o % S == rank
        if (kSorted) {
            // ---- phase 1 (sorted): the slice is bucketed by area class, the chunk's candidates are ordered by area class;
            //      a warp takes 4 class-adjacent candidates (warp-uniform registers) and its LANES walk only the kept
            //      boxes whose area can reach them: min(area) / max(area) >= thr is necessary for IoU >= thr, so
            //      everything outside that range of classes is skipped exactly.
            // (a) re-bucket the slice when boxes were appended by the previous chunk (counting sort into the other
            //     buffer) and (b) order the chunk's candidates by class (counting sort of <= 256 positions; positions
            //     past the end of the list are marked suppressed right away) -- the two sorts share their barriers
            const bool resort = ns > ns_sorted;
            for (int e = tid; e < kKeyCap; e += kThreads) { sm->shist[e] = 0u; sm->chist[e] = 0u; }
            __syncthreads();
            if (resort)
                for (int i = tid; i < ns; i += kThreads) atomicAdd(&sm->shist[strip_of(kbox[i], karea[i])], 1u);
            int cst = kKeys + 1;
            if (tid < kChunk) {
                if (base + tid < cnt) cst = strip_of(sm->cbox[tid], sm->carea[tid]);
                else atomicOr(&sm->acc[par][tid >> 5], 1u << (tid & 31));
                atomicAdd(&sm->chist[cst], 1u);
            }
            __syncthreads();
            if (warp < 2 && (warp == 1 || resort)) {  // warp 0: slice keys, warp 1: candidate keys (kKeyPer keys per lane)
                unsigned int* hist = warp == 0 ? sm->shist : sm->chist;
                unsigned int* cursor = warp == 0 ? sm->scursor : sm->ccursor;
                unsigned int h[kKeyPer], t3 = 0;
#pragma unroll
                for (int q = 0; q < kKeyPer; q += 4) {  // independent 16-byte loads (entries past kKeys + 1 are zero)
                    const uint4 v = *reinterpret_cast<const uint4*>(hist + lane * kKeyPer + q);
                    h[q] = v.x; h[q + 1] = v.y; h[q + 2] = v.z; h[q + 3] = v.w;
                    t3 += v.x + v.y + v.z + v.w;
                }
                unsigned int inc = t3;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                unsigned int run = inc - t3;
#pragma unroll
                for (int q = 0; q < kKeyPer; q += 4) {
                    uint4 v;
                    v.x = run; run += h[q];
                    v.y = run; run += h[q + 1];
                    v.z = run; run += h[q + 2];
                    v.w = run; run += h[q + 3];
                    *reinterpret_cast<uint4*>(cursor + lane * kKeyPer + q) = v;
                    if (warp == 0) {
                        const int e = lane * kKeyPer + q;
                        if (e + 3 <= kKeys + 1) {
                            *reinterpret_cast<uint2*>(sm->sstart + e) =
                                make_uint2((v.x & 0xffffu) | (v.y << 16), (v.z & 0xffffu) | (v.w << 16));
                        } else {
                            if (e <= kKeys + 1) sm->sstart[e] = (unsigned short)v.x;
                            if (e + 1 <= kKeys + 1) sm->sstart[e + 1] = (unsigned short)v.y;
                            if (e + 2 <= kKeys + 1) sm->sstart[e + 2] = (unsigned short)v.z;
                        }
                    }
                }
            }
            __syncthreads();
            if (resort) {
                for (int i = tid; i < ns; i += kThreads) {
                    const float4 b = kbox[i];
                    const float a = karea[i];
                    const unsigned int pos = atomicAdd(&sm->scursor[strip_of(b, a)], 1u);
                    kbox_alt[pos] = b;
                    karea_alt[pos] = a;
                }
                float4* tb = kbox; kbox = kbox_alt; kbox_alt = tb;
                float* ta = karea; karea = karea_alt; karea_alt = ta;
                ns_sorted = ns;
            }
            if (tid < kChunk) sm->cord[atomicAdd(&sm->ccursor[cst], 1u)] = (unsigned short)tid;
            __syncthreads();
            FRR_TICK(10);  // bucketing time (reported separately, not part of the phase-1 slot)
            // (c) kSub threads per candidate (candidates in class order, so the lanes of a warp walk ranges of similar
            //     length): each thread screens its candidate against every kSub-th kept box of the candidate's admissible
            //     classes and of the always-tested class, remembers one hit, and confirms it with the exact test.  (The
            //     earlier form -- a warp per 4 candidates, lanes over kept boxes -- spent half of its issue slots on
            //     warp-uniform bookkeeping: after the area cut a group's range is only ~4 warp trips long.  Dealing
            //     fixed-size pieces of the ranges to the threads through a prefix sum balances better but was measured
            //     1.7x slower: the per-item search and the virtual-range indexing cost more than the imbalance.)
            {
                constexpr int kSub = kThreads / kChunk;
                const int sub = tid % kSub;
                const int cpos = sm->cord[tid / kSub];
                if (base + cpos < cnt) {
                    const float4 cbx = sm->cbox[cpos];
                    const float ca = sm->carea[cpos];
                    int pk = -1;
                    int c_lo = 0, c_hi = -1, x_lo = 0, x_hi = 0;  // admissible keys: classes c_lo..c_hi, x bins x_lo..x_hi
                    int lo2 = 0, hi2 = ns;                        // second segment: everything (NaN area) / always-tested
                    if (ca == ca) {
                        const float a = box_area(cbx);
                        c_lo = strip_of_area(a * thr.alo);
                        c_hi = strip_of_area(a * thr.ahi);
                        const float cx = 0.5f * (cbx.x + cbx.z);
                        const float rx = thr.fx * (cbx.z - cbx.x) * 1.0001f + 2.0e-6f;
                        x_lo = xbin_of(cx - rx);
                        x_hi = xbin_of(cx + rx);
                        lo2 = sm->sstart[kKeys];
                        hi2 = sm->sstart[kKeys + 1];
                    }
                    // the bounds of all admissible classes are fetched before the first walk (independent loads): with one
                    // dependent sstart -> kbox chain per class the walks were latency bound
                    constexpr int kMaxCls = 6;  // [alo, ahi] spans 2 * log2(1 / thr) * 4 + 1 classes: 5.1 at thr 0.7
                    int lo_c[kMaxCls], hi_c[kMaxCls];
#pragma unroll
                    for (int q = 0; q < kMaxCls; ++q) {
                        const int c = min(c_lo + q, kStrips - 1);
                        lo_c[q] = sm->sstart[c * kXBins + x_lo];
                        hi_c[q] = (c_lo + q <= c_hi) ? (int)sm->sstart[c * kXBins + x_hi + 1] : 0;
                    }
#pragma unroll
                    for (int q = 0; q < kMaxCls; ++q) {
#pragma unroll 2
                        for (int k = lo_c[q] + sub; k < hi_c[q]; k += kSub)
                            if (suppress_screen<true>(kbox[k], karea[k], cbx, ca)) pk = k;
                    }
                    for (int c = c_lo + kMaxCls; c <= c_hi; ++c) {  // thresholds below 0.6: more classes
                        const int lo = sm->sstart[c * kXBins + x_lo], hi = sm->sstart[c * kXBins + x_hi + 1];
                        for (int k = lo + sub; k < hi; k += kSub)
                            if (suppress_screen<true>(kbox[k], karea[k], cbx, ca)) pk = k;
                    }
                    for (int k = lo2 + sub; k < hi2; k += kSub)
                        if (suppress_screen<true>(kbox[k], karea[k], cbx, ca)) pk = k;
                    if (pk >= 0) {
                        bool r = suppress_exact(kbox[pk], cbx, thr.up);
                        if (!r) {  // the screen hit was not confirmed by the exact test (rare): exact walk of the own share
                            for (int c = c_lo; c <= c_hi && !r; ++c) {
                                const int lo = sm->sstart[c * kXBins + x_lo], hi = sm->sstart[c * kXBins + x_hi + 1];
                                for (int k = lo + sub; k < hi && !r; k += kSub) r = suppress_exact(kbox[k], cbx, thr.up);
                            }
                            for (int k = lo2 + sub; k < hi2 && !r; k += kSub) r = suppress_exact(kbox[k], cbx, thr.up);
                        }
                        if (r) atomicOr(&sm->acc[par][cpos >> 5], 1u << (cpos & 31));
                    }
                }
            }
        } else {
        // ---- phase 1: 4 candidates per thread vs this CTA's slice of the kept list ----------------------
        {
            float4 cb[kTile];
            float ca[kTile];
            bool has[kTile], sup[kTile];
#pragma unroll
            for (int j = 0; j < kTile; ++j) {
                cb[j] = sm->cbox[u + j * kGroup];
                ca[j] = sm->carea[u + j * kGroup];
                has[j] = base + u + j * kGroup < cnt;
                sup[j] = !has[j];
            }
            if (kFast) {
                // Screen every kept box of the slice; remember only the FIRST one that passes the screen and
                // evaluate it exactly after the loop (all lanes together).  The screen is tight (2^-19), so a
                // pass is almost always a true suppression; the rare lane whose exact test fails rescans.
                // The slice part is walked from its END so that the plain conditional move below leaves the SMALLEST
                // passing k (one ALU op less per pair than "first hit wins").
                int pk[kTile];
#pragma unroll
                for (int j = 0; j < kTile; ++j) pk[j] = -1;
                const int cntp = (ns > part) ? (ns - part + kParts - 1) / kParts : 0;
#pragma unroll 2
                for (int k = part + (cntp - 1) * kParts; k >= part; k -= kParts) {
                    const float4 kb = kbox[k];
                    const float ka = karea[k];
#pragma unroll
                    for (int j = 0; j < kTile; ++j)
                        if (suppress_screen<kUnit>(kb, ka, cb[j], ca[j])) pk[j] = k;
                }
#pragma unroll
                for (int j = 0; j < kTile; ++j) {
                    if (has[j] && pk[j] >= 0) {
                        sup[j] = suppress_exact(kbox[pk[j]], cb[j], thr.up);
                        for (int k2 = pk[j] + kParts; k2 < ns && !sup[j]; k2 += kParts)
                            if (suppress_screen<kUnit>(kbox[k2], karea[k2], cb[j], ca[j]))
                                sup[j] = suppress_exact(kbox[k2], cb[j], thr.up);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kTile; ++j)
                    for (int k = part; k < ns && !sup[j]; k += kParts) sup[j] = suppress_exact(kbox[k], cb[j], thr.up);
            }
            // warp covers candidates u in [32 * (warp % kGW), +32) + kGroup j  ->  word (warp % kGW) + kGW j
            constexpr int kGW = kGroup / 32;
#pragma unroll
            for (int j = 0; j < kTile; ++j) {
                const unsigned int wj = __ballot_sync(0xffffffffu, sup[j]);
                if (lane == 0 && wj != 0u) atomicOr(&sm->acc[par][(warp % kGW) + kGW * j], wj);
            }
        }
        }
        __syncthreads();
        // publish this CTA's 8 words to every CTA of the cluster (distributed shared memory)
        if (tid < kChunkWords * S) {
            const int wd = tid % kChunkWords, dstr = tid / kChunkWords;
            unsigned int* dst = cluster.map_shared_rank(&sm->supp[par][rank][wd], dstr);
            *dst = sm->acc[par][wd];
        }
        FRR_TICK(DBG_P1);
        cluster.sync();
        FRR_TICK(DBG_SYNC);

        // ---- phase 2: combine the S words per 32 candidates, compact survivors (first kChunk threads) ---
        bool alive = false;
        unsigned int am = 0;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float fa = 0.f;
        if (tid < kChunk) {
            const unsigned int v = (lane < S) ? sm->supp[par][lane][warp] : 0u;
            const unsigned int all = __reduce_or_sync(0xffffffffu, v);
            alive = !((all >> lane) & 1u);
            am = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) sm->warp_cnt[warp] = __popc(am);
            if (tid < kChunkWords) {
                sm->kept_w[0][tid] = 0;
                sm->removed_w[0][tid] = 0;
                sm->acc[par ^ 1][tid] = 0;  // for the next chunk
            }
            bx = sm->cbox[tid];
            fa = sm->carea[tid];
        }
        __syncthreads();
        int s = 0;
        {
            int soff = 0;
#pragma unroll
            for (int w2 = 0; w2 < kChunkWords; ++w2) {
                const int cc = (int)sm->warp_cnt[w2];
                if (w2 < warp) soff += cc;
                s += cc;
            }
            if (alive) {
                const int si = soff + __popc(am & ((1u << lane) - 1u));
                sm->sbox[si] = bx;
                sm->sarea[si] = fa;
                sm->ssrc[si] = (short)tid;
            }
        }
        __syncthreads();
        FRR_TICK(DBG_P2);
        if (prof) dbg[DBG_SURV] += s;

        // ---- phase 3: predecessor masks among survivors.  3a: warp item = (row, 32-column word), screen only (two
        //      words per trip for ILP); 3b: the few set bits are refined with the exact test, one (row, word) entry
        //      per thread, so the exact path never diverges a busy warp ----------------------------------------
        for (int row = warp; row < s; row += kWarps) {
            const float4 rb = sm->sbox[row];
            const float ra = sm->sarea[row];
            for (int q = 0; q * 32 < row; q += 2) {
                const int j0 = q * 32 + lane, j1 = j0 + 32;  // j1 <= 255
                bool h0 = false, h1 = false;
                if (kFast) {
                    h0 = (j0 < row) && suppress_screen<kUnit>(sm->sbox[j0], sm->sarea[j0], rb, ra);
                    h1 = (j1 < row) && suppress_screen<kUnit>(sm->sbox[j1], sm->sarea[j1], rb, ra);
                } else {
                    if (j0 < row) h0 = suppress_exact(sm->sbox[j0], rb, thr.up);
                    if (j1 < row) h1 = suppress_exact(sm->sbox[j1], rb, thr.up);
                }
                const unsigned int b0 = __ballot_sync(0xffffffffu, h0);
                const unsigned int b1 = __ballot_sync(0xffffffffu, h1);
                if (lane == 0) {
                    sm->pred[row][q] = b0;
                    sm->pred[row][q + 1] = b1;
                }
            }
        }
        if (kFast) {
            __syncthreads();
            for (int e = tid; e < s * kChunkWords; e += kThreads) {
                const int row = e / kChunkWords, q = e - row * kChunkWords;
                if (q * 32 < row) {
                    unsigned int m = sm->pred[row][q], keepm = m;
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        if (!suppress_exact(sm->sbox[q * 32 + bit], sm->sbox[row], thr.up)) keepm &= ~(1u << bit);
                    }
                    sm->pred[row][q] = keepm;
                }
            }
        }
        __syncthreads();
        FRR_TICK(DBG_P3);

        // ---- phase 4: parallel fix-point over the survivors (first kChunk threads, named barrier 1).
        //      survivor i is kept once every predecessor that suppresses it is removed; removed as soon
        //      as one such predecessor is kept.  State words are double buffered: one barrier / round. ----
        int fin = 0;
        if (tid < kChunk) {
            int state = (tid < s) ? 0 : 3;  // 0 undecided, 1 kept, 2 removed, 3 n/a
            unsigned int p[kChunkWords];
#pragma unroll
            for (int q = 0; q < kChunkWords; ++q) p[q] = (state == 0 && q * 32 < tid) ? sm->pred[tid][q] : 0u;
            int cur = 0;
            for (;;) {
                if (state == 0) {
                    bool hit_kept = false, pending = false;
#pragma unroll
                    for (int q = 0; q < kChunkWords; ++q) {
                        const unsigned int kw = sm->kept_w[cur][q], rw = sm->removed_w[cur][q];
                        hit_kept |= (p[q] & kw) != 0u;
                        pending |= (p[q] & ~(kw | rw)) != 0u;
                    }
                    if (hit_kept) state = 2;
                    else if (!pending) state = 1;
                }
                const unsigned int km = __ballot_sync(0xffffffffu, state == 1);
                const unsigned int rm = __ballot_sync(0xffffffffu, state == 2);
                if (lane == 0) {
                    sm->kept_w[cur ^ 1][warp] = km;
                    sm->removed_w[cur ^ 1][warp] = rm;
                }
                cur ^= 1;
                if (prof) dbg[DBG_ITERS] += 1;
                if (!bar_or(1, kChunk, state == 0)) break;
            }
            fin = cur;
        }
        // every thread needs the final buffer index: it is uniform over the first kChunk threads
        fin = __syncthreads_or(fin);
        FRR_TICK(DBG_P4);

        // ---- phase 5: append kept survivors --------------------------------------------------------------
        {
            int koff = 0, ktot = 0;
#pragma unroll
            for (int w2 = 0; w2 < kChunkWords; ++w2) {
                const int cc = __popc(sm->kept_w[fin][w2]);
                if (w2 < warp) koff += cc;
                ktot += cc;
            }
            if (tid < kChunk) {
                const unsigned int kw = sm->kept_w[fin][warp];
                if ((kw >> lane) & 1u) {
                    const int o = nk + koff + __popc(kw & ((1u << lane) - 1u));
                    if (o < max_keep) {
                        const float4 kb = sm->sbox[tid];
                        if (o % S == rank) {
                            kbox[o / S] = kb;
                            karea[o / S] = sm->sarea[tid];
                        }
                        if (rank == 0) {
                            ikeep[o] = base + (int)sm->ssrc[tid];
                            if (iout) iout[o] = kb;
                        }
                    }
                }
            }
            nk = min(nk + ktot, max_keep);
        }
        __syncthreads();
        FRR_TICK(DBG_P5);
    }
#undef FRR_TICK

    if (rank == 0) {
        for (int o = nk + tid; o < max_keep; o += kThreads) {
            ikeep[o] = -1;
            if (iout) iout[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid == 0) keep_count[img] = nk;
    }
}

static NmsThr make_thr(double thr) {
    NmsThr t;
    float f = (float)thr;
    if (isnan(thr)) {
        t.up = NAN;  // nothing is ever > NaN
    } else {
        if (!((double)f > thr)) f = nextafterf(f, INFINITY);
        t.up = f;
    }
    t.fast = (thr >= 1.0e-6) && isfinite(thr) && (t.up < 1.0e30f) ? 1 : 0;
    const double u = (double)t.up;
    t.c2 = t.fast ? (float)(u / (1.0 + u) * (1.0 - 1.9073486328125e-06)) : 0.f;
    // IoU >= thr needs min(area) / max(area) >= thr; thr is lowered by 2^-17 relative to cover the fp32 roundings of the
    // exact IoU (<= 2^-20, see strip_of_area) and of the two products below
    const double tl = thr * (1.0 - 7.62939453125e-06);
    // ... and |cx_a - cx_b| <= max(1 - thr, (1 - thr) / (2 thr)) * w of EITHER box
    t.fx = t.fast ? (float)(fmax(1.0 - tl, (1.0 - tl) / (2.0 * tl)) * (1.0 + 1.0e-6)) : 0.f;
    t.alo = t.fast ? (float)(tl * (1.0 - 1.0e-6)) : 0.f;
    t.ahi = t.fast ? (float)(1.0 / tl * (1.0 + 1.0e-6)) : 0.f;
    return t;
}

static size_t nms_smem_bytes(int slice_cap, bool sorted) {
    return ((sizeof(NmsSmem) + 15) & ~(size_t)15) + (size_t)slice_cap * (sizeof(float4) + sizeof(float)) * (sorted ? 2 : 1);
}

// Launch geometry and kernel variant for a problem: shared by nms_launch and frr_nms_variant (tests / bench assert
// through the latter that the variant they mean to exercise is the one that runs).
struct NmsPick {
    int S, threads, slice_cap, sorted;
    size_t smem;
    NmsThr thr;
};

static int nms_pick(int B, int n, double iou_thr, int max_keep, int cluster_size, int threads, int unit_boxes, NmsPick* out) {
    FRR_CHECK_ARG(B >= 0 && n >= 0 && max_keep >= 0, "frr_nms_sorted: bad sizes B=%d n=%d max_keep=%d", B, n, max_keep);
    FRR_CHECK_ARG(threads == 0 || threads == 256 || threads == 512 || threads == 1024,
                  "frr_nms_sorted: threads %d not in {0,256,512,1024}", threads);
    const int kcap = max_keep < n ? max_keep : n;  // most boxes that can ever be kept
    int S = cluster_size;
    if (S == 0) {
        // auto: fill the machine.  148 SMs / B images, rounded down to a power of two, capped at 16.
        int per = num_sms() / (B > 0 ? B : 1);
        S = 1;
        while (S * 2 <= per && S < 16) S *= 2;
    }
    FRR_CHECK_ARG(S == 1 || S == 2 || S == 4 || S == 8 || S == 16, "frr_nms_sorted: cluster_size %d not in {1,2,4,8,16}", S);
    // a CTA whose slice of the kept list is short (one image spread over 16 CTAs) is bound by its per-chunk barriers:
    // 16 warps resolve them faster than 32 (130 vs 134 us for 12000 -> 2000 boxes)
    if (threads == 0) threads = (kcap / S < 192) ? 512 : 1024;
    if (threads < kChunk) threads = kChunk;  // the first kChunk threads own one candidate each
    const NmsThr thr = make_thr(iou_thr);
    // The class-sorted phase 1 pays ~10 k cycles of bucketing per chunk: it wins once a CTA's slice of the kept
    // list is large (batched launches with 1-2 CTAs per image), not for a single image spread over 16 CTAs.
    const bool sorted = thr.fast && unit_boxes && (kcap / S >= 384);
    // grow the cluster until a slice of the kept list fits in shared memory
    const size_t limit = 227 * 1024;
    while (nms_smem_bytes((kcap + S - 1) / S + 1, sorted) > limit && S < 16) S *= 2;
    const int slice_cap = (kcap + S - 1) / S + 1;
    const size_t smem = nms_smem_bytes(slice_cap, sorted);
    FRR_CHECK_ARG(smem <= limit, "frr_nms_sorted: max_keep=%d does not fit the kept list in shared memory", max_keep);
    out->S = S;
    out->threads = threads;
    out->slice_cap = slice_cap;
    out->sorted = sorted ? 1 : 0;
    out->smem = smem;
    out->thr = thr;
    return FRR_OK;
}

int nms_launch(const float* boxes, const int32_t* counts, int B, int n, double iou_thr, int max_keep,
                      int32_t* keep, int32_t* keep_count, float* out_boxes, int cluster_size, int threads,
                      long long* dbg, int unit_boxes, frr_stream_t stream, const int32_t* gather_idx, int src_n) {
    FRR_CHECK_ARG(keep && keep_count, "frr_nms_sorted: null output");
    FRR_CHECK_ARG(n == 0 || (boxes && aligned16(boxes)), "frr_nms_sorted: boxes must be non-null, 16-byte aligned");
    FRR_CHECK_ARG(out_boxes == nullptr || aligned16(out_boxes), "frr_nms_sorted: out_boxes must be 16-byte aligned");
    NmsPick pk;
    {
        const int rc = nms_pick(B, n, iou_thr, max_keep, cluster_size, threads, unit_boxes, &pk);
        if (rc) return rc;
    }
    if (B == 0) return FRR_OK;
    const int S = pk.S, slice_cap = pk.slice_cap;
    threads = pk.threads;
    const bool sorted = pk.sorted != 0;
    const size_t smem = pk.smem, limit = 227 * 1024;
    const NmsThr thr = pk.thr;

    using kern_t = void (*)(const float4*, const int32_t*, int, int, int, NmsThr, int32_t*, int32_t*, float4*, long long*,
                            const int32_t*, int);
    kern_t kern = nullptr;
#define FRR_NMS_PICK(F, U, SO)                                                                      \
    (threads < kChunk * 2 ? nms_keeplist_kernel<kChunk, F, U, SO>                                            \
                    : threads == 512 ? nms_keeplist_kernel<512, F, U, SO> : nms_keeplist_kernel<1024, F, U, SO>)
    if (sorted) kern = FRR_NMS_PICK(true, true, true);
    else if (thr.fast && unit_boxes) kern = FRR_NMS_PICK(true, true, false);
    else if (thr.fast) kern = FRR_NMS_PICK(true, false, false);
    else kern = FRR_NMS_PICK(false, false, false);
#undef FRR_NMS_PICK
    FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    if (S > 8) FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * S), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FRR_CUDA(cudaLaunchKernelEx(&cfg, kern, (const float4*)boxes, counts, n, max_keep, slice_cap, thr, keep, keep_count,
                                (float4*)out_boxes, dbg, gather_idx, src_n));
    count_launch();
    FRR_CHECK_LAUNCH("nms_keeplist_kernel");
    return FRR_OK;
}

}  // namespace frr

extern "C" int frr_nms_sorted(const float* boxes, const int32_t* counts, int B, int n, double iou_thr, int max_keep,
                              int32_t* keep, int32_t* keep_count, float* out_boxes, int cluster_size,
                              frr_stream_t stream) {
    return frr::nms_launch(boxes, counts, B, n, iou_thr, max_keep, keep, keep_count, out_boxes, cluster_size, 0, nullptr,
                           0, stream, nullptr, 0);
}

extern "C" int frr_nms_sorted_tuned(const float* boxes, const int32_t* counts, int B, int n, double iou_thr, int max_keep,
                                    int32_t* keep, int32_t* keep_count, float* out_boxes, int cluster_size, int threads,
                                    int64_t* dbg_cycles, int unit_boxes, frr_stream_t stream) {
    return frr::nms_launch(boxes, counts, B, n, iou_thr, max_keep, keep, keep_count, out_boxes, cluster_size, threads,
                           (long long*)dbg_cycles, unit_boxes, stream, nullptr, 0);
}

// Which kernel variant / launch geometry frr_nms_sorted* picks for a problem (host only, launches nothing):
// out[0] = CTAs per image (cluster size), out[1] = threads per CTA, out[2] = variant (0 exact only, 1 screened,
// 2 screened + unit range, 3 unit range + class / x-bin bucketed kept slice), out[3] = dynamic shared memory bytes.
extern "C" int frr_nms_variant(int B, int n, double iou_thr, int max_keep, int cluster_size, int threads, int unit_boxes,
                               int32_t* out4) {
    FRR_CHECK_ARG(out4 != nullptr, "frr_nms_variant: null output");
    frr::NmsPick pk;
    const int rc = frr::nms_pick(B, n, iou_thr, max_keep, cluster_size, threads, unit_boxes, &pk);
    if (rc) return rc;
    out4[0] = pk.S;
    out4[1] = pk.threads;
    out4[2] = pk.sorted ? 3 : (pk.thr.fast ? (unit_boxes ? 2 : 1) : 0);
    out4[3] = (int32_t)pk.smem;
    return FRR_OK;
}

// Same as frr_nms_sorted_tuned, but the score order is given as indices: candidate i of image b is
// boxes_src[b][order[b][i]] (boxes_src [B,src_n,4], order int32 [B,n], e.g. the out_idx of frr_topk_desc).
extern "C" int frr_nms_sorted_indirect(const float* boxes_src, int src_n, const int32_t* order, const int32_t* counts, int B,
                                       int n, double iou_thr, int max_keep, int32_t* keep, int32_t* keep_count,
                                       float* out_boxes, int cluster_size, int unit_boxes, frr_stream_t stream) {
    FRR_CHECK_ARG(order != nullptr && src_n >= 0, "frr_nms_sorted_indirect: order must be given");
    return frr::nms_launch(boxes_src, counts, B, n, iou_thr, max_keep, keep, keep_count, out_boxes, cluster_size, 0, nullptr,
                           unit_boxes, stream, order, src_n);
}
