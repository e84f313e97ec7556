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
//   only (~5.5 M pair tests instead of the 72 M of the full upper triangle).
//
// Bit-exactness vs the CPU kernel: IoU = inter / ((area_i + area_j) - inter) in fp32 with IEEE
// division and no FMA contraction; suppress when (double)iou > thr  <=>  iou >= thr_up, where
// thr_up is the smallest fp32 strictly above thr.  Almost every pair is decided without the
// division by a guarded product test (inter vs thr_up*(1 -+ 2^-20)*union); only pairs inside
// the guard band take the exact division.  Degenerate boxes carry a NaN "fast area" so that
// they always fall through to the exact path.
#include <cooperative_groups.h>
#include <math.h>

#include "frr_common.cuh"

namespace cg = cooperative_groups;

namespace frr {

constexpr int kNmsThreads = 256;  // = chunk size
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kMaxCluster = 16;

struct NmsThr {
    float up;    // smallest fp32 with (double)up > thr
    float c_lo;  // up * (1 - 2^-20): below -> certainly not suppressed
    float c_hi;  // up * (1 + 2^-20): above -> certainly suppressed
    int fast;    // guarded product test usable (thr > 0 and finite)
};

__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// area used by the fast path: NaN for boxes that are not well formed (forces the exact path)
__device__ __forceinline__ float fast_area(const float4& b) {
    const float a = box_area(b);
    const bool ok = (b.z >= b.x) && (b.w >= b.y) && (a <= 3.0e38f);
    return ok ? a : __int_as_float(0x7fc00000);
}

__device__ __noinline__ bool suppress_exact(const float4& a, const float4& b, float up) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(box_area(a), box_area(b)), inter);
    const float ovr = __fdiv_rn(inter, uni);
    return ovr >= up;  // false for NaN, as (double)NaN > thr
}

// a = earlier (kept) box, b = candidate; fa/fb = fast areas.  Symmetric in (a,b).
__device__ __forceinline__ bool suppresses(const float4& a, float fa, const float4& b, float fb, const NmsThr& t) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(fa, fb), inter);
    // !(inter < lo) is also true when uni is NaN (degenerate box) -> exact path
    if (!(inter < __fmul_rn(t.c_lo, uni))) {
        if (inter > __fmul_rn(t.c_hi, uni)) return true;
        return suppress_exact(a, b, t.up);
    }
    return false;
}

struct NmsSmem {
    unsigned int supp[2][kMaxCluster][kNmsWarps];  // [parity][source rank][warp]: suppressed-by-slice words
    unsigned int warp_cnt[kNmsWarps];
    unsigned int kept_w[kNmsWarps];     // fix-point state: survivor bit sets (<= 256 survivors)
    unsigned int removed_w[kNmsWarps];
    unsigned int pred[kNmsThreads][kNmsWarps];  // pred[i][q]: bit j of word q set if survivor 32q+j (< i) suppresses i
    float4 sbox[kNmsThreads];                   // survivor boxes (compacted)
    float sarea[kNmsThreads];
    short ssrc[kNmsThreads];                    // survivor -> position inside the chunk
    int undecided;
};

template <bool kFast>
__global__ void __launch_bounds__(kNmsThreads)
    nms_keeplist_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ counts, int n, int max_keep,
                        int slice_cap, NmsThr thr, int32_t* __restrict__ keep, int32_t* __restrict__ keep_count,
                        float4* __restrict__ out_boxes) {
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int img = blockIdx.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem* sm = reinterpret_cast<NmsSmem*>(smem_raw);
    float4* kbox = reinterpret_cast<float4*>(smem_raw + ((sizeof(NmsSmem) + 15) & ~(size_t)15));  // [slice_cap]
    float* karea = reinterpret_cast<float*>(kbox + slice_cap);                                      // [slice_cap]

    const int cnt = counts ? min(counts[img], n) : n;
    const float4* ib = boxes + (size_t)img * n;
    int32_t* ikeep = keep + (size_t)img * max_keep;
    float4* iout = out_boxes ? out_boxes + (size_t)img * max_keep : nullptr;

    int nk = 0;  // kept so far (identical in every CTA of the cluster)
    int par = 0;
    for (int base = 0; base < cnt && nk < max_keep; base += kNmsThreads, par ^= 1) {
        const int i = base + tid;
        const bool has = i < cnt;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) bx = ib[i];
        const float fa = fast_area(bx);

        // ---- phase 1: candidate vs this CTA's slice of the kept list -----------------------------
        const int ns = (nk - rank + S - 1) / S;  // kept ordinals o with o % S == rank
        bool sup = !has;
        {
            int k = 0;
            for (; k + 4 <= ns; k += 4) {
                bool s0, s1, s2, s3;
                if (kFast) {
                    s0 = suppresses(kbox[k], karea[k], bx, fa, thr);
                    s1 = suppresses(kbox[k + 1], karea[k + 1], bx, fa, thr);
                    s2 = suppresses(kbox[k + 2], karea[k + 2], bx, fa, thr);
                    s3 = suppresses(kbox[k + 3], karea[k + 3], bx, fa, thr);
                } else {
                    s0 = suppress_exact(kbox[k], bx, thr.up);
                    s1 = suppress_exact(kbox[k + 1], bx, thr.up);
                    s2 = suppress_exact(kbox[k + 2], bx, thr.up);
                    s3 = suppress_exact(kbox[k + 3], bx, thr.up);
                }
                sup |= (s0 | s1) | (s2 | s3);
                if ((k & 31) == 28 && __all_sync(0xffffffffu, sup)) { k = ns; break; }
            }
            for (; k < ns; ++k)
                sup |= kFast ? suppresses(kbox[k], karea[k], bx, fa, thr) : suppress_exact(kbox[k], bx, thr.up);
        }
        const unsigned int word = __ballot_sync(0xffffffffu, sup);
        if (lane < S) {
            unsigned int* dst = cluster.map_shared_rank(&sm->supp[par][rank][warp], lane);
            *dst = word;
        }
        cluster.sync();

        // ---- phase 2: combine, compact survivors ---------------------------------------------------
        unsigned int all = 0;
        for (int r = 0; r < S; ++r) all |= sm->supp[par][r][warp];
        const bool alive = !((all >> lane) & 1u);
        const unsigned int am = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) sm->warp_cnt[warp] = __popc(am);
        if (tid < kNmsWarps) { sm->kept_w[tid] = 0; sm->removed_w[tid] = 0; }
        __syncthreads();
        int soff = 0, s = 0;
#pragma unroll
        for (int w2 = 0; w2 < kNmsWarps; ++w2) {
            const int c = (int)sm->warp_cnt[w2];
            if (w2 < warp) soff += c;
            s += c;
        }
        if (alive) {
            const int si = soff + __popc(am & ((1u << lane) - 1u));
            sm->sbox[si] = bx;
            sm->sarea[si] = fa;
            sm->ssrc[si] = (short)tid;
        }
        __syncthreads();

        // ---- phase 3: predecessor masks among survivors (warp item = (row i, 32-column word q)) -----
        const int nw = (s + 31) >> 5;
        for (int it = warp; it < s * nw; it += kNmsWarps) {
            const int row = it / nw, q = it - row * nw;
            if (q * 32 >= row) continue;  // only columns j < row
            const int j = q * 32 + lane;
            bool hit = false;
            if (j < row) {
                hit = kFast ? suppresses(sm->sbox[j], sm->sarea[j], sm->sbox[row], sm->sarea[row], thr)
                            : suppress_exact(sm->sbox[j], sm->sbox[row], thr.up);
            }
            const unsigned int m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) sm->pred[row][q] = m;
        }
        __syncthreads();

        // ---- phase 4: parallel fix-point.  survivor i is kept when every predecessor that
        //      suppresses it is removed; removed as soon as one such predecessor is kept. --------------
        {
            const bool mine = tid < s;
            const int myw = (tid >> 5) + 1;  // words that can hold predecessors of `tid`
            int state = mine ? 0 : 3;        // 0 undecided, 1 kept, 2 removed, 3 n/a
            for (;;) {
                if (state == 0) {
                    bool hit_kept = false, pending = false;
                    for (int q = 0; q < myw && q < nw; ++q) {
                        if (q * 32 >= tid) break;
                        const unsigned int p = sm->pred[tid][q];
                        hit_kept |= (p & sm->kept_w[q]) != 0u;
                        pending |= (p & ~(sm->kept_w[q] | sm->removed_w[q])) != 0u;
                    }
                    if (hit_kept) state = 2;
                    else if (!pending) state = 1;
                }
                __syncthreads();  // all reads of kept_w/removed_w done before they are updated
                const unsigned int km = __ballot_sync(0xffffffffu, state == 1);
                const unsigned int rm = __ballot_sync(0xffffffffu, state == 2);
                if (lane == 0) { sm->kept_w[warp] = km; sm->removed_w[warp] = rm; }
                if (__syncthreads_or(state == 0) == 0) break;
            }
        }
        // kept_w now final (visible after the barrier inside __syncthreads_or)

        // ---- phase 5: append kept survivors ---------------------------------------------------------
        {
            int koff = 0, ktot = 0;
#pragma unroll
            for (int w2 = 0; w2 < kNmsWarps; ++w2) {
                const int c = __popc(sm->kept_w[w2]);
                if (w2 < warp) koff += c;
                ktot += c;
            }
            const unsigned int kw = sm->kept_w[warp];
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
            nk = min(nk + ktot, max_keep);
        }
        __syncthreads();
    }

    if (rank == 0) {
        for (int o = nk + tid; o < max_keep; o += kNmsThreads) {
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
    t.fast = (thr > 0.0) && isfinite(thr) && (t.up < 1.0e30f) ? 1 : 0;
    t.c_lo = t.up * (1.0f - 9.5367431640625e-07f);
    t.c_hi = t.up * (1.0f + 9.5367431640625e-07f);
    return t;
}

static size_t nms_smem_bytes(int slice_cap) {
    return ((sizeof(NmsSmem) + 15) & ~(size_t)15) + (size_t)slice_cap * (sizeof(float4) + sizeof(float));
}

}  // namespace frr

extern "C" int frr_nms_sorted(const float* boxes, const int32_t* counts, int B, int n, double iou_thr, int max_keep,
                              int32_t* keep, int32_t* keep_count, float* out_boxes, int cluster_size,
                              frr_stream_t stream) {
    using namespace frr;
    FRR_CHECK_ARG(keep && keep_count, "frr_nms_sorted: null output");
    FRR_CHECK_ARG(B >= 0 && n >= 0 && max_keep >= 0, "frr_nms_sorted: bad sizes B=%d n=%d max_keep=%d", B, n, max_keep);
    FRR_CHECK_ARG(n == 0 || (boxes && aligned16(boxes)), "frr_nms_sorted: boxes must be non-null, 16-byte aligned");
    FRR_CHECK_ARG(out_boxes == nullptr || aligned16(out_boxes), "frr_nms_sorted: out_boxes must be 16-byte aligned");
    if (B == 0) return FRR_OK;
    if (max_keep > n) {
        // keep buffers are [B,max_keep]; more than n can never be kept, but the stride stays max_keep
    }
    const int kcap = max_keep < n ? max_keep : n;  // most boxes that can ever be kept
    int S = cluster_size;
    if (S == 0) {
        // auto: fill the machine.  148 SMs / B images, rounded down to a power of two, capped at 8 (portable).
        int per = num_sms() / (B > 0 ? B : 1);
        S = 1;
        while (S * 2 <= per && S < 8) S *= 2;
    }
    FRR_CHECK_ARG(S == 1 || S == 2 || S == 4 || S == 8 || S == 16, "frr_nms_sorted: cluster_size %d not in {1,2,4,8,16}", S);
    // grow the cluster until a slice of the kept list fits in shared memory
    const size_t limit = 227 * 1024;
    while (nms_smem_bytes((kcap + S - 1) / S + 1) > limit && S < 16) S *= 2;
    const int slice_cap = (kcap + S - 1) / S + 1;
    const size_t smem = nms_smem_bytes(slice_cap);
    FRR_CHECK_ARG(smem <= limit, "frr_nms_sorted: max_keep=%d does not fit the kept list in shared memory", max_keep);

    const NmsThr thr = make_thr(iou_thr);
    auto kern = thr.fast ? nms_keeplist_kernel<true> : nms_keeplist_kernel<false>;
    FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    if (S > 8) FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * S), 1, 1);
    cfg.blockDim = dim3(kNmsThreads, 1, 1);
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
                                (float4*)out_boxes));
    count_launch();
    FRR_CHECK_LAUNCH("nms_keeplist_kernel");
    return FRR_OK;
}
