// N1 for the RPN proposal layer (models/model.py:53-55: nms(roi, softmax, 0.7), keep[:2000|300]) on boxes clamped to
// [0,1] (models/model.py:34): greedy NMS in LARGE chunks with every pair test restricted to admissible buckets.
//
// One thread-block cluster per image; every CTA holds the WHOLE kept list in shared memory, the candidates of a chunk
// (up to 2048, in score order) are dealt round-robin to the CTAs.  Per chunk:
//   A-C  the kept list is re-bucketed (counting sort by key = (area class, x bin), see nms_common.cuh) and every CTA
//        orders its own candidates by key, so the lanes of a warp walk ranges of similar length;
//   D    screen: a candidate walks only the kept boxes whose keys are admissible for it -- IoU >= thr needs an area
//        ratio in [thr, 1/thr] and an x-centre distance <= fx * width, both exact necessary conditions -- with the
//        division-free screen; one screen hit is confirmed with the exact torchvision arithmetic.  Dead candidates
//        are flagged in every CTA through distributed shared memory (one cluster barrier);
//   E-G  the chunk's survivors are bucketed the same way, in a DETERMINISTIC order (key, then score position: an
//        unordered counting-sort scatter followed by a rank inside the bucket), so every CTA builds the same array;
//   H    suppression pairs among the survivors: the survivor at bucketed position p walks only the admissible
//        positions ABOVE p (both pruning rules are symmetric in the pair, so every pair that can suppress is met
//        exactly once, from its lower position; positions are dealt round-robin to the CTAs).  Screen hits are parked
//        in a per-thread ring and confirmed with the exact test after the walk; a confirmed pair becomes an EDGE
//        (later, earlier in score order), appended to the CTA's edge region and mirrored into every CTA (second
//        cluster barrier).  A survivor whose edges do not fit is flagged and decided by a walk of its own instead;
//   I    fix-point over the edges (every CTA, redundantly): a survivor is removed as soon as one predecessor is kept
//        and kept once all its predecessors are removed -- edge pass + node pass per round, 3-6 rounds on RPN boxes;
//   J    the kept survivors are appended in score order (block scan), the walk stops at max_keep.
// Against the keep-list kernel of nms.cu (chunks of 256, ~10 barriers each): 3-4 chunks instead of 18-25, i.e. 3-4
// latency chains per image, and 0.4-0.6 M pair tests per image instead of 4.6 M.  The keep list is the one of
// torchvision's CPU kernel bit for bit: every decision is taken by suppress_exact(); screen and buckets only skip
// pairs that provably cannot suppress.  tests/nms_model.py restates this algorithm on the CPU.
#include <cooperative_groups.h>

#include <atomic>

#include "nms_common.cuh"

namespace cg = cooperative_groups;

namespace frr {

constexpr int kBkChunk = 2048;     // most candidates per chunk (shared-memory arrays are sized for it)
constexpr int kBkListCap = 2048;   // kept-list capacity (max_keep <= 2047)
constexpr int kBkEdgeCap = 8192;   // edges per chunk over the whole cluster (each CTA owns kBkEdgeCap / S)
constexpr int kBkBins = 3;         // x bins of a walk whose bounds are prefetched (more: tail loop)
constexpr int kBkUpBins = 2;       // same for the upward walk (bins above the own one)
constexpr int kBkHits = 4;         // screen hits a thread parks per walk before it falls back to the slow walk

// lanes_screen 0 = automatic: 2 lanes per candidate when an image has 1-2 CTAs (batched launches: 146 vs 153 us with one
// CTA per image, tools/nms_s1_sweep.py), 4 in the large clusters of the single-image form (81 vs 85 us)
static std::atomic<int> g_bk_kc0{1024}, g_bk_kcmax{2048}, g_bk_nsub1{0}, g_bk_nsub2{4};

struct BkSmem {
    unsigned int lhist[kKeyCap];  // kept list: key histogram, then scatter cursors
    unsigned int chist[kKeyCap];  // this CTA's candidates
    unsigned int shist[kKeyCap];  // the chunk's survivors
    unsigned short lstart[kKeyCap];  // first list position of key s; [kKeys + 1] = list length
    unsigned short sstart[kKeyCap];  // same for the bucketed survivors
    unsigned int edges[kBkEdgeCap];  // (later << 16) | earlier, chunk-local positions; region r belongs to CTA r
    unsigned int ecount[kMaxCluster];
    unsigned int ecursor;
    unsigned int work[2];  // dynamic dealing of the items of phases D and H
    unsigned int warp_tmp[32];
    float4 cbox[kBkChunk];  // the chunk's candidates, score order
    float csa[kBkChunk];    // c2-scaled screening areas (NaN = always take the exact path)
    unsigned short ckey[kBkChunk];
    unsigned short cord[kBkChunk];  // this CTA's candidates ordered by key
    unsigned short sidx[kBkChunk];  // bucketed survivor -> chunk position
    unsigned char cstate[2][kBkChunk];  // [chunk parity]: 0 undecided, 1 kept, 2 removed
    unsigned char pend[kBkChunk];
    unsigned char ovf[kBkChunk];
    unsigned int hitbuf[1024 * 4];  // per-thread ring of screen hits (phase H)
};

enum {
    BK_CHUNKS = 0, BK_LOAD, BK_SCAN, BK_SCATTER, BK_SCREEN, BK_SYNC1, BK_SURV_SORT, BK_PRED, BK_SYNC2, BK_FIX, BK_APPEND,
    BK_ROUNDS, BK_SURVIVORS, BK_EDGES, BK_VISITED, BK_OVF
};

// One warp: in-place exclusive scan of hist[0 .. kKeyCap) (kKeyPer consecutive entries per lane), optionally also
// written as 16-bit start offsets.  Entries past kKeys + 1 are zero.
__device__ __forceinline__ void bk_scan_keys(unsigned int* hist, unsigned short* start, int lane) {
    unsigned int h[kKeyPer], t3 = 0;
#pragma unroll
    for (int q = 0; q < kKeyPer; q += 4) {
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
        *reinterpret_cast<uint4*>(hist + lane * kKeyPer + q) = v;
        if (start)
            *reinterpret_cast<uint2*>(start + lane * kKeyPer + q) =
                make_uint2((v.x & 0xffffu) | (v.y << 16), (v.z & 0xffffu) | (v.w << 16));
    }
}

// Bucket key = (x bin, area class), x bin major: for one x bin the admissible area classes of a box are ONE contiguous
// run of the bucketed array, so a walk is a handful of long ranges (one per admissible x bin) instead of one short range
// per class.  Boxes without a usable screening area get the extra key kKeys (always tested, last bucket).
__device__ __forceinline__ int bk_key(const float4& b, float sa) {
    return (sa != sa) ? kKeys : xbin_of(0.5f * (b.x + b.z)) * kStrips + strip_of_area(box_area(b));
}

struct BkAdm {
    int c_lo, c_hi, x_lo, x_hi;
};
// Admissible keys of a box: area classes of [alo * A, ahi * A] (IoU <= min / max area) and x bins of cx -+ fx * w.
__device__ __forceinline__ BkAdm bk_admissible(const float4& cbx, const NmsThr& thr) {
    BkAdm r;
    const float a = box_area(cbx);
    r.c_lo = strip_of_area(a * thr.alo);
    r.c_hi = strip_of_area(a * thr.ahi);
    const float cx = 0.5f * (cbx.x + cbx.z);
    const float rx = thr.fx * (cbx.z - cbx.x) * 1.0001f + 2.0e-6f;
    r.x_lo = xbin_of(cx - rx);
    r.x_hi = xbin_of(cx + rx);
    return r;
}

// Calls f(k) for every position k of a bucketed array (start offsets `start`) whose key is admissible for the box
// (cbx, screening area ca); this thread takes every nsub-th position of each range, starting at sub.
template <class F>
__device__ __forceinline__ void bk_walk(const unsigned short* __restrict__ start, const float4& cbx, float ca,
                                        const NmsThr& thr, int sub, int nsub, F&& f) {
    if (ca == ca) {
        const BkAdm ad = bk_admissible(cbx, thr);
        // the bounds of the first bins are fetched before the first walk (independent loads)
        int lo_b[kBkBins], hi_b[kBkBins];
#pragma unroll
        for (int q = 0; q < kBkBins; ++q) {
            const int xb = min(ad.x_lo + q, kXBins - 1);
            lo_b[q] = start[xb * kStrips + ad.c_lo];
            hi_b[q] = (ad.x_lo + q <= ad.x_hi) ? (int)start[xb * kStrips + ad.c_hi + 1] : 0;
        }
        const int lo2 = start[kKeys], hi2 = start[kKeys + 1];  // boxes that must always be tested
#pragma unroll
        for (int q = 0; q < kBkBins; ++q) {
#pragma unroll 2
            for (int k = lo_b[q] + sub; k < hi_b[q]; k += nsub) f(k);
        }
        for (int xb = ad.x_lo + kBkBins; xb <= ad.x_hi; ++xb) {  // wide boxes / low thresholds: more bins
            const int lo = start[xb * kStrips + ad.c_lo], hi = start[xb * kStrips + ad.c_hi + 1];
            for (int k = lo + sub; k < hi; k += nsub) f(k);
        }
        for (int k = lo2 + sub; k < hi2; k += nsub) f(k);
    } else {  // no usable screening area: everything
        const int hi = start[kKeys + 1];
        for (int k = sub; k < hi; k += nsub) f(k);
    }
}

// Upward half of bk_walk for the element at position p of the bucketed array (its key: own_key): positions above p
// only.  Both pruning rules are symmetric in a pair, so a pair that can suppress is found from its lower position.
template <class F>
__device__ __forceinline__ void bk_walk_up(const unsigned short* __restrict__ start, int p, int own_key, const float4& cbx,
                                           float ca, const NmsThr& thr, int sub, int nsub, F&& f) {
    const int total = start[kKeys + 1];
    if (ca == ca) {
        const BkAdm ad = bk_admissible(cbx, thr);
        const int xb0 = own_key / kStrips;
        const int hi0 = start[xb0 * kStrips + ad.c_hi + 1];  // own bin: the rest of the own bucket and the classes up to c_hi
        int lo_b[kBkUpBins], hi_b[kBkUpBins];
#pragma unroll
        for (int q = 0; q < kBkUpBins; ++q) {
            const int xb = min(xb0 + 1 + q, kXBins - 1);
            lo_b[q] = start[xb * kStrips + ad.c_lo];
            hi_b[q] = (xb0 + 1 + q <= ad.x_hi) ? (int)start[xb * kStrips + ad.c_hi + 1] : 0;
        }
        const int lo2 = start[kKeys];
#pragma unroll 2
        for (int k = p + 1 + sub; k < hi0; k += nsub) f(k);
#pragma unroll
        for (int q = 0; q < kBkUpBins; ++q) {
#pragma unroll 2
            for (int k = lo_b[q] + sub; k < hi_b[q]; k += nsub) f(k);
        }
        for (int xb = xb0 + 1 + kBkUpBins; xb <= ad.x_hi; ++xb) {
            const int lo = start[xb * kStrips + ad.c_lo], hi = start[xb * kStrips + ad.c_hi + 1];
            for (int k = lo + sub; k < hi; k += nsub) f(k);
        }
        for (int k = lo2 + sub; k < total; k += nsub) f(k);
    } else {  // the always-tested bucket is the last one
        for (int k = p + 1 + sub; k < total; k += nsub) f(k);
    }
}

// Same positions and the same dealing to the sub-threads as bk_walk, one compact copy of the code (rare paths: kept
// out of the hot loops' footprint).  f returns true to stop.
template <class F>
__device__ __noinline__ void bk_walk_slow(const unsigned short* __restrict__ start, float4 cbx, float ca, NmsThr thr,
                                          int sub, int nsub, F f) {
    int lo2 = 0;
    const int hi2 = start[kKeys + 1];
    if (ca == ca) {
        const BkAdm ad = bk_admissible(cbx, thr);
        for (int xb = ad.x_lo; xb <= ad.x_hi; ++xb) {
            const int lo = start[xb * kStrips + ad.c_lo], hi = start[xb * kStrips + ad.c_hi + 1];
            for (int k = lo + sub; k < hi; k += nsub)
                if (f(k)) return;
        }
        lo2 = start[kKeys];
    }
    for (int k = lo2 + sub; k < hi2; k += nsub)
        if (f(k)) return;
}

// bk_walk_up's positions with the same dealing to the sub-threads, one compact copy (a thread whose hit ring
// overflowed re-walks exactly ITS share).
template <class F>
__device__ __noinline__ void bk_walk_up_slow(const unsigned short* __restrict__ start, int p, int own_key, float4 cbx, float ca,
                                             NmsThr thr, int sub, int nsub, F f) {
    const int total = start[kKeys + 1];
    if (ca == ca) {
        const BkAdm ad = bk_admissible(cbx, thr);
        const int xb0 = own_key / kStrips;
        const int hi0 = start[xb0 * kStrips + ad.c_hi + 1];
        for (int k = p + 1 + sub; k < hi0; k += nsub) f(k);
        for (int xb = xb0 + 1; xb <= ad.x_hi; ++xb) {
            const int lo = start[xb * kStrips + ad.c_lo], hi = start[xb * kStrips + ad.c_hi + 1];
            for (int k = lo + sub; k < hi; k += nsub) f(k);
        }
        for (int k = start[kKeys] + sub; k < total; k += nsub) f(k);
    } else {
        for (int k = p + 1 + sub; k < total; k += nsub) f(k);
    }
}

// A confirmed suppression pair of the chunk: the later box (chunk position i) has the earlier one (j) as predecessor.
// Appended to this CTA's edge region (mirrored into the other CTAs after the phase); if the region is full the later
// box is flagged and decides from a walk of its own in the fix-point.
__device__ __forceinline__ void bk_push_edge(BkSmem* sm, int i, int j, int rank, int ecap) {
    const unsigned int e = atomicAdd(&sm->ecursor, 1u);
    if (e < (unsigned int)ecap) sm->edges[rank * ecap + e] = ((unsigned int)i << 16) | (unsigned int)j;
    else sm->ovf[i] = 1;
}

// A warp takes the next 32 / nsub items of a phase (dynamic dealing: walks differ in length by more than 10x).
__device__ __forceinline__ int bk_grab(unsigned int* cursor, int per_warp, int lane) {
    int q0 = 0;
    if (lane == 0) q0 = (int)atomicAdd(cursor, (unsigned int)per_warp);
    return __shfl_sync(0xffffffffu, q0, 0);
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads)
    nms_bucket_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ counts, int n, int max_keep, NmsThr thr,
                      int32_t* __restrict__ keep, int32_t* __restrict__ keep_count, float4* __restrict__ out_boxes,
                      long long* __restrict__ dbg, const int32_t* __restrict__ gather_idx, int src_n, int kc0, int kc_max, int nsub1,
                      int nsub2) {
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int img = blockIdx.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    BkSmem* sm = reinterpret_cast<BkSmem*>(smem_raw);
    float4* bbox[2];
    float* barea[2];
    bbox[0] = reinterpret_cast<float4*>(smem_raw + ((sizeof(BkSmem) + 15) & ~(size_t)15));
    bbox[1] = bbox[0] + kBkListCap;
    barea[0] = reinterpret_cast<float*>(bbox[1] + kBkListCap);
    barea[1] = barea[0] + kBkListCap;

    const int cnt = counts ? min(counts[img], n) : n;
    const float4* ib = boxes + (size_t)img * (gather_idx ? src_n : n);
    const int32_t* gi = gather_idx ? gather_idx + (size_t)img * n : nullptr;
    int32_t* ikeep = keep + (size_t)img * max_keep;
    float4* iout = out_boxes ? out_boxes + (size_t)img * max_keep : nullptr;
    const int ecap = kBkEdgeCap / S;

    const bool prof = (dbg != nullptr) && blockIdx.x == 0 && tid == 0;
    long long t0 = 0;
#define BK_TICK(slot)                   \
    if (prof) {                         \
        const long long t1 = clock64(); \
        dbg[slot] += t1 - t0;           \
        t0 = t1;                        \
    }

    for (int e = tid; e < kKeyCap; e += kThreads) {
        sm->lhist[e] = 0u;
        sm->chist[e] = 0u;
        sm->shist[e] = 0u;
    }
    for (int i = tid; i < kBkChunk; i += kThreads) sm->cstate[0][i] = 0;
    if (tid == 0) {
        sm->ecursor = 0u;
        sm->work[0] = 0u;
        sm->work[1] = 0u;
    }
    // (one CTA per image: a block barrier does; the hardware cluster barrier costs ~1 us even for a cluster of one)
    auto cluster_sync = [&]() { if (S == 1) __syncthreads(); else cluster.sync(); };
    cluster_sync();  // every CTA's shared memory is initialised before a peer writes into it

    int nk = 0;   // kept so far (identical in every CTA of the cluster)
    int cur = 0;  // buffer that holds the kept list
    int par = 0;  // chunk parity (cstate double buffer)
    int kc = kc0;
    for (int base = 0; base < cnt && nk < max_keep; par ^= 1) {
        const int m = min(kc, cnt - base);
        if (prof) { t0 = clock64(); dbg[BK_CHUNKS] += 1; dbg[BK_VISITED] += m; }
        unsigned char* st = sm->cstate[par];
        // ---- A: stage the chunk (screening areas, keys), histograms of the list keys and of the own candidates' keys
        for (int i = tid; i < m; i += kThreads) {
            const float4 b = gi ? ib[gi[base + i]] : ib[base + i];
            const float sa = screen_area(b, thr.c2);
            const int key = bk_key(b, sa);
            sm->cbox[i] = b;
            sm->csa[i] = sa;
            sm->ckey[i] = (unsigned short)key;
            sm->pend[i] = 0;
            sm->ovf[i] = 0;
            if (i % S == rank) atomicAdd(&sm->chist[key], 1u);
        }
        for (int k = tid; k < nk; k += kThreads) atomicAdd(&sm->lhist[bk_key(bbox[cur][k], barea[cur][k])], 1u);
        __syncthreads();
        BK_TICK(BK_LOAD);
        // ---- B: exclusive scans (warp 0: list, warp 1: own candidates)
        if (warp == 0) bk_scan_keys(sm->lhist, sm->lstart, lane);
        else if (warp == 1) bk_scan_keys(sm->chist, nullptr, lane);
        __syncthreads();
        BK_TICK(BK_SCAN);
        // ---- C: scatter the list into the other buffer, the own candidates into cord
        for (int k = tid; k < nk; k += kThreads) {
            const float4 b = bbox[cur][k];
            const float a = barea[cur][k];
            const unsigned int pos = atomicAdd(&sm->lhist[bk_key(b, a)], 1u);
            bbox[cur ^ 1][pos] = b;
            barea[cur ^ 1][pos] = a;
        }
        for (int i = tid; i < m; i += kThreads)
            if (i % S == rank) sm->cord[atomicAdd(&sm->chist[sm->ckey[i]], 1u)] = (unsigned short)i;
        cur ^= 1;
        __syncthreads();
        BK_TICK(BK_SCATTER);
        const float4* lb = bbox[cur];
        const float* la = barea[cur];
        // ---- D: screen the own candidates against the admissible part of the kept list: nsub1 lanes per candidate,
        //      a warp takes 32 / nsub1 candidates (adjacent in key order: walks of similar length) at a time
        const int n_own = (m - rank + S - 1) / S;
        if (nk > 0) {
            const int per_warp = 32 / nsub1, sub = lane % nsub1;
            for (;;) {
                const int q = bk_grab(&sm->work[0], per_warp, lane) + lane / nsub1;
                if (q - lane / nsub1 >= n_own) break;
                if (q >= n_own) continue;
                const int i = sm->cord[q];
                const float4 cbx = sm->cbox[i];
                const float ca = sm->csa[i];
                int pk = -1;
                bk_walk(sm->lstart, cbx, ca, thr, sub, nsub1, [&](int k) {
                    if (suppress_screen<true>(lb[k], la[k], cbx, ca)) pk = k;
                });
                if (pk >= 0) {
                    bool r = suppress_exact(lb[pk], cbx, thr.up);
                    if (!r) {  // the screen hit was not confirmed by the exact test (rare): exact walk of the own share
                        const float up = thr.up;
                        bool* rp = &r;
                        bk_walk_slow(sm->lstart, cbx, ca, thr, sub, nsub1, [=](int k) {
                            if (suppress_screen<true>(lb[k], la[k], cbx, ca) && suppress_exact(lb[k], cbx, up)) *rp = true;
                            return *rp;
                        });
                    }
                    if (r) {
                        st[i] = 2;
                        for (int d = 1; d < S; ++d) *cluster.map_shared_rank(&st[i], (rank + d) % S) = 2;
                    }
                }
            }
        }
        BK_TICK(BK_SCREEN);
        cluster_sync();
        BK_TICK(BK_SYNC1);
        // ---- E: histogram of the survivors' keys; reset what the next chunk's phases A-D accumulate into
        for (int i = tid; i < kBkChunk; i += kThreads) sm->cstate[par ^ 1][i] = 0;
        for (int e = tid; e < kKeyCap; e += kThreads) {
            sm->lhist[e] = 0u;
            sm->chist[e] = 0u;
        }
        for (int i = tid; i < m; i += kThreads)
            if (st[i] == 0) atomicAdd(&sm->shist[sm->ckey[i]], 1u);
        __syncthreads();
        // ---- F: scan
        if (warp == 0) bk_scan_keys(sm->shist, sm->sstart, lane);
        __syncthreads();
        // ---- G: bucket the survivors into the free list buffer, ordered by (key, chunk position): an unordered
        //      scatter of the positions, then every entry ranks itself inside its bucket -- the array is the same in
        //      every CTA of the cluster, which phase H relies on to deal the pairs out
        float4* sb = bbox[cur ^ 1];
        float* sa_ = barea[cur ^ 1];
        unsigned short* tmp = sm->cord;  // the own-candidate order is not needed any more
        for (int i = tid; i < m; i += kThreads)
            if (st[i] == 0) tmp[atomicAdd(&sm->shist[sm->ckey[i]], 1u)] = (unsigned short)i;
        __syncthreads();
        const int ns = sm->sstart[kKeys + 1];
        for (int t = tid; t < ns; t += kThreads) {
            const int i = tmp[t];
            const int key = sm->ckey[i];
            const int lo = sm->sstart[key], hi = sm->sstart[key + 1];
            int r = 0;
            for (int u = lo; u < hi; ++u) r += (tmp[u] < i) ? 1 : 0;
            const int p = lo + r;
            sb[p] = sm->cbox[i];
            sa_[p] = sm->csa[i];
            sm->sidx[p] = (unsigned short)i;
        }
        __syncthreads();
        BK_TICK(BK_SURV_SORT);
        if (prof) dbg[BK_SURVIVORS] += ns;
        // ---- H: suppression pairs among the survivors, every pair from its lower bucketed position (nsub2 lanes per
        //      survivor, dealt like phase D)
        {
            const int ns_own = (ns - rank + S - 1) / S;
            const int per_warp = 32 / nsub2, sub = lane % nsub2;
            unsigned int* hb = sm->hitbuf + tid * kBkHits;
            for (;;) {
                const int q = bk_grab(&sm->work[1], per_warp, lane) + lane / nsub2;
                if (q - lane / nsub2 >= ns_own) break;
                if (q >= ns_own) continue;
                const int p = q * S + rank;
                const int i = sm->sidx[p];
                const float4 cbx = sb[p];
                const float ca = sa_[p];
                int nh = 0;
                bk_walk_up(sm->sstart, p, sm->ckey[i], cbx, ca, thr, sub, nsub2, [&](int k) {
                    if (suppress_screen<true>(sb[k], sa_[k], cbx, ca)) {
                        hb[nh & (kBkHits - 1)] = (unsigned int)k;
                        ++nh;
                    }
                });
                if (nh > kBkHits) {  // more screen hits than the ring holds: walk the share again, confirming as they come
                    const float up = thr.up;
                    const unsigned short* sidx = sm->sidx;
                    BkSmem* smp = sm;
                    bk_walk_up_slow(sm->sstart, p, sm->ckey[i], cbx, ca, thr, sub, nsub2, [=](int k) {
                        if (suppress_screen<true>(sb[k], sa_[k], cbx, ca) && suppress_exact(sb[k], cbx, up)) {
                            const int j = sidx[k];
                            bk_push_edge(smp, max(i, j), min(i, j), rank, ecap);
                        }
                    });
                } else {
                    for (int h = 0; h < nh; ++h) {
                        const int k = (int)hb[h];
                        if (suppress_exact(sb[k], cbx, thr.up)) {
                            const int j = sm->sidx[k];
                            bk_push_edge(sm, max(i, j), min(i, j), rank, ecap);
                        }
                    }
                }
            }
        }
        __syncthreads();
        {   // mirror this CTA's edges (and, if its region overflowed, the flags) into the other CTAs of the cluster
            const unsigned int raw = sm->ecursor;
            const int ne = (int)min(raw, (unsigned int)ecap);
            for (int d = 1; d < S; ++d) {
                const int dst = (rank + d) % S;
                unsigned int* re = cluster.map_shared_rank(&sm->edges[rank * ecap], dst);
                for (int e = tid; e < ne; e += kThreads) re[e] = sm->edges[rank * ecap + e];
                if (raw > (unsigned int)ecap) {
                    unsigned char* ro = cluster.map_shared_rank(&sm->ovf[0], dst);
                    for (int i = tid; i < m; i += kThreads)
                        if (sm->ovf[i]) ro[i] = 1;
                }
            }
            if (tid < S) *cluster.map_shared_rank(&sm->ecount[rank], tid) = (unsigned int)ne;
        }
        BK_TICK(BK_PRED);
        cluster_sync();
        BK_TICK(BK_SYNC2);
        // ---- I: fix-point (every CTA on all survivors and all edges)
        for (;;) {
            if (prof) dbg[BK_ROUNDS] += 1;
            {   // a thread works on one region: regions fill evenly (positions are dealt round-robin)
                const int r = tid % S;
                const int ne = (int)sm->ecount[r];
                for (int e = tid / S; e < ne; e += kThreads / S) {
                    const unsigned int v = sm->edges[r * ecap + e];
                    const int i = (int)(v >> 16), j = (int)(v & 0xffffu);
                    if (st[i] == 0) {
                        const int sj = st[j];
                        if (sj == 1) st[i] = 2;
                        else if (sj == 0) sm->pend[i] = 1;
                    }
                }
            }
            __syncthreads();
            bool any = false;
            for (int i = tid; i < m; i += kThreads) {
                if (st[i] != 0) continue;
                if (sm->ovf[i]) {  // edges incomplete: decide from a walk over all admissible predecessors
                    bool hit_kept = false, pending = false;
                    bool* hk = &hit_kept;
                    bool* pd = &pending;
                    const float4 cbx = sm->cbox[i];
                    const float ca = sm->csa[i];
                    const float up = thr.up;
                    const unsigned short* sidx = sm->sidx;
                    const unsigned char* stc = st;
                    bk_walk_slow(sm->sstart, cbx, ca, thr, 0, 1, [=](int p) {
                        const int j = sidx[p];
                        if (j < i) {
                            const int sj = stc[j];
                            if (sj != 2 && suppress_screen<true>(sb[p], sa_[p], cbx, ca) && suppress_exact(sb[p], cbx, up)) {
                                if (sj == 1) *hk = true;
                                else *pd = true;
                            }
                        }
                        return *hk;
                    });
                    sm->pend[i] = 0;
                    if (hit_kept) st[i] = 2;
                    else if (!pending) st[i] = 1;
                    else any = true;
                } else if (sm->pend[i]) {
                    sm->pend[i] = 0;
                    any = true;
                } else {
                    st[i] = 1;
                }
            }
            if (!__syncthreads_or(any)) break;
        }
        BK_TICK(BK_FIX);
        if (prof)
            for (int r = 0; r < S; ++r) dbg[BK_EDGES] += sm->ecount[r];
        // ---- J: append the kept survivors in score order
        int kept_chunk;
        {
            constexpr int kPer = kBkChunk / kThreads;  // consecutive candidates per thread
            const int i0 = tid * kPer;
            unsigned int c = 0;
#pragma unroll
            for (int q = 0; q < kPer; ++q) c += (i0 + q < m && st[i0 + q] == 1) ? 1u : 0u;
            unsigned int total;
            int o = nk + (int)block_exclusive_scan(c, sm->warp_tmp, &total);
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                const int i = i0 + q;
                if (i < m && st[i] == 1) {
                    if (o < max_keep) {
                        const float4 kb = sm->cbox[i];
                        bbox[cur][o] = kb;
                        barea[cur][o] = sm->csa[i];
                        if (rank == 0) {
                            ikeep[o] = base + i;
                            if (iout) iout[o] = kb;
                        }
                    }
                    ++o;
                }
            }
            kept_chunk = (int)total;
            nk = min(nk + kept_chunk, max_keep);
        }
        for (int e = tid; e < kKeyCap; e += kThreads) sm->shist[e] = 0u;
        if (tid == 0) {
            sm->ecursor = 0u;
            sm->work[0] = 0u;
            sm->work[1] = 0u;
        }
        base += m;
        // next chunk: enough candidates for the missing keeps at the rate of this chunk, plus half
        {
            const long long need = max_keep - nk;
            long long next = (need * m * 3) / (2 * (long long)max(kept_chunk, 1)) + 1;
            next = (next + 255) & ~255LL;
            kc = (int)min((long long)kc_max, max(256LL, next));
        }
        __syncthreads();
        BK_TICK(BK_APPEND);
    }
#undef BK_TICK

    if (rank == 0) {
        for (int o = nk + tid; o < max_keep; o += kThreads) {
            ikeep[o] = -1;
            if (iout) iout[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid == 0) keep_count[img] = nk;
    }
}

bool nms_bucket_eligible(int n, int max_keep, const NmsThr& thr, int unit_boxes) {
    const int kcap = max_keep < n ? max_keep : n;
    return thr.fast && unit_boxes && kcap < kBkListCap && n >= 1;
}

size_t nms_bucket_smem_bytes() {
    return ((sizeof(BkSmem) + 15) & ~(size_t)15) + (size_t)2 * kBkListCap * (sizeof(float4) + sizeof(float));
}

int nms_bucket_launch(const float* boxes, const int32_t* counts, int B, int n, const NmsThr& thr, int max_keep, int32_t* keep,
                      int32_t* keep_count, float* out_boxes, int S, int threads, long long* dbg, frr_stream_t stream,
                      const int32_t* gather_idx, int src_n) {
    using kern_t = void (*)(const float4*, const int32_t*, int, int, NmsThr, int32_t*, int32_t*, float4*, long long*,
                            const int32_t*, int, int, int, int, int);
    kern_t kern = threads <= 256 ? nms_bucket_kernel<256> : threads == 512 ? nms_bucket_kernel<512> : nms_bucket_kernel<1024>;
    if (threads < 256) threads = 256;
    const size_t smem = nms_bucket_smem_bytes();
    FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (S > 8) FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // first chunk: small when few boxes are wanted (test mode: 300), at most kc0
    int kc_max = g_bk_kcmax.load(std::memory_order_relaxed);
    int kc0 = g_bk_kc0.load(std::memory_order_relaxed);
    kc_max = kc_max < 256 ? 256 : (kc_max > kBkChunk ? kBkChunk : (kc_max & ~255));
    int want = ((max_keep + max_keep / 2) + 255) & ~255;
    if (want < 256) want = 256;
    kc0 = kc0 < 256 ? 256 : (kc0 & ~255);
    if (kc0 > want) kc0 = want;
    if (kc0 > kc_max) kc0 = kc_max;

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
    FRR_CUDA(cudaLaunchKernelEx(&cfg, kern, (const float4*)boxes, counts, n, max_keep, thr, keep, keep_count,
                                (float4*)out_boxes, dbg, gather_idx, src_n, kc0, kc_max,
                                g_bk_nsub1.load(std::memory_order_relaxed) ? g_bk_nsub1.load(std::memory_order_relaxed)
                                                                           : (S <= 2 ? 2 : 4),
                                g_bk_nsub2.load(std::memory_order_relaxed)));
    count_launch();
    FRR_CHECK_LAUNCH("nms_bucket_kernel");
    return FRR_OK;
}

}  // namespace frr

// Developer knob (profiling tools): sizes of the first chunk and of the largest chunk of the bucketed kernel
// (multiples of 256, <= 2048) and the lanes per candidate in the screen / pair phases (1, 2, 4, ... 32; 0 keeps the
// current value).  Process-wide; results never depend on them.
extern "C" int frr_nms_bucket_tune(int first_chunk, int max_chunk, int lanes_screen, int lanes_pairs) {
    FRR_CHECK_ARG(first_chunk >= 256 && max_chunk >= 256 && first_chunk <= frr::kBkChunk && max_chunk <= frr::kBkChunk,
                  "frr_nms_bucket_tune: chunk sizes must lie in [256, %d]", frr::kBkChunk);
    auto pow2 = [](int v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; };
    FRR_CHECK_ARG((lanes_screen == 0 || pow2(lanes_screen)) && (lanes_pairs == 0 || pow2(lanes_pairs)),
                  "frr_nms_bucket_tune: lanes per candidate must be a power of two <= 32");
    frr::g_bk_kc0.store(first_chunk);
    frr::g_bk_kcmax.store(max_chunk);
    frr::g_bk_nsub1.store(lanes_screen);  // (0: back to automatic)
    if (lanes_pairs) frr::g_bk_nsub2.store(lanes_pairs);
    return FRR_OK;
}
