// T1/T3 sampling on the device: the reference draws its 256-anchor / 128-RoI samples with torch.randperm on the HOST
// mt19937 generator (models/model.py:149,155,228,235; models/new_model.py:169-177,328-343), and how many numbers each
// draw consumes depends on the data.  To keep the sampled indices bit-exact under torch.manual_seed WITHOUT the
// device -> host -> device round trip, the generator state itself lives in device memory:
//
//   sample_stream_kernel (one CTA): reads the candidate counts of the assign kernels, lays out the draws exactly in the
//     reference's order (per image: RPN positives, RPN negatives, Fast R-CNN positives, Fast R-CNN negatives), then
//     walks the mt19937 stream block by block (624-word twists, one barrier each, double buffered) and keeps only
//     the tempered words the samplers will look at -- torch's randperm(n) is a forward Fisher-Yates shuffle
//     (r[i] <-> r[i + random() % (n - i)], i = 0 .. n-2), so the first `take` entries of the permutation need only the
//     first `take` draws although the call consumes n - 1 of them.  The advanced state is written back.
//   sample_apply_kernel (one CTA per draw): replays those `take` Fisher-Yates steps on an implicit identity array
//     (a small shared-memory hash map holds the displaced entries) and applies the result: RPN labels outside the
//     kept sample become -1 (:228-236), the Fast R-CNN selection rows `sel` / `sel_n` are written for the finalize kernel.
//
// Bit-exactness: integer arithmetic only; pinned against torch.randperm in tests (KAT-5) and against the host path.
#include "frr_common.cuh"

namespace frr {

constexpr int kMtN = 624;
constexpr int kMtM = 397;
constexpr int kD = kMtN - kMtM;  // 227: distance below which the twist has no dependency on its own output
constexpr int kStreamThreads = 640;  // one state word per thread in the twist
constexpr int kApplyThreads = 512;
constexpr int kMaxTake = 512;    // entries of a permutation that are ever looked at (256 RPN, 128 / 512 Fast R-CNN)
constexpr int kHashSize = 2048;  // >= 4 x kMaxTake
constexpr int kJobsSmem = 1024;  // the stream kernel walks its job table in shared memory: <= 256 images per launch

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// job j = 4 * image + kind (0 RPN positives, 1 RPN negatives, 2 Fast R-CNN positives, 3 Fast R-CNN negatives)
//   x = n      length of the permutation (0: the reference does not draw here)
//   y = need   draws the sampler reads = min(take, n - 1)
//   z = take   leading entries of the permutation that are used = min(wanted, n)
//   w = off    position of the job's first draw in the stream (relative to the state at entry)
__global__ void __launch_bounds__(kStreamThreads)
    sample_stream_kernel(const int32_t* __restrict__ rpn_counts, const int32_t* __restrict__ fr_counts, int B, int rpn_batch,
                         int rpn_max_pos, int fr_batch, int fr_max_pos, uint32_t* __restrict__ state, int4* __restrict__ jobs,
                         uint32_t* __restrict__ draws, int S) {
    __shared__ uint32_t mt_buf[2][kMtN];
    __shared__ unsigned int warp_tmp[32];
    __shared__ int4 sjobs[kJobsSmem];  // the job table is walked once per twist: keep it out of the global-load latency
    const int tid = threadIdx.x;

    // ---- job table (one image per thread, images in order) --------------------------------------
    unsigned int run = 0;
    for (int base = 0; base < B; base += kStreamThreads) {
        const int b = base + tid;
        int n[4] = {0, 0, 0, 0}, want[4] = {0, 0, 0, 0};
        unsigned int cons = 0;
        if (b < B) {
            if (rpn_counts) {  // models/model.py:225-236
                const int np = rpn_counts[2 * b], nn = rpn_counts[2 * b + 1];
                if (np > rpn_max_pos) { n[0] = np; want[0] = rpn_max_pos; }
                if (nn > rpn_batch - np) { n[1] = nn; want[1] = rpn_batch - min(np, rpn_max_pos); }
            }
            if (fr_counts) {  // models/model.py:144-156: both permutations are always drawn
                const int np = fr_counts[2 * b], nn = fr_counts[2 * b + 1];
                const int kp = min(np, fr_max_pos);
                n[2] = np; want[2] = kp;
                n[3] = nn; want[3] = fr_batch - kp;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) cons += (unsigned int)max(n[q] - 1, 0);
        }
        unsigned int total;
        unsigned int off = run + block_exclusive_scan(cons, warp_tmp, &total);
        run += total;
        if (b < B) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int take = min(max(want[q], 0), n[q]);
                const int4 jb = make_int4(n[q], min(take, max(n[q] - 1, 0)), take, (int)off);
                jobs[4 * b + q] = jb;
                sjobs[4 * b + q] = jb;
                off += (unsigned int)max(n[q] - 1, 0);
            }
        }
        __syncthreads();
    }
    const unsigned int T = run;  // numbers this batch consumes
    const int J = 4 * B;

    // ---- the stream ---------------------------------------------------------------------------------
    for (int i = tid; i < kMtN; i += kStreamThreads) mt_buf[0][i] = state[i];
    int pos = (int)state[kMtN];  // next unread word; kMtN = exhausted (twist first)
    __syncthreads();
    int cur = 0;
    unsigned int consumed = 0;
    // Every thread walks the (uniform) job table, so the walk must cost nothing in the common case of a 624-word block
    // nobody reads: j = next job with draws still to emit, [lo_j, hi_j) = its stream range (lo_j = "never" when done).
    int j = 0;
    unsigned int lo_j = 0xffffffffu, hi_j = 0;
    auto next_job = [&]() {
        while (j < J && sjobs[j].y == 0) ++j;
        if (j < J) { lo_j = (unsigned int)sjobs[j].w; hi_j = lo_j + (unsigned int)sjobs[j].y; }
        else lo_j = 0xffffffffu;
    };
    next_job();
    while (consumed < T) {
        if (pos >= kMtN) {
            // One barrier per twist: the sequential algorithm updates word i from the OLD words i, i + 1 and the word
            // 397 ahead (already NEW for i >= 227).  Substituting the new words by their own definitions expresses every
            // new word through old words only (the recurrence is XOR-linear in the far operand):
            //   i < 227        : o[i+397]                                                     ^ f(i)
            //   227 <= i < 454 : o[i+170] ^ f(i-227)                                          ^ f(i)
            //   454 <= i < 623 : o[i-57]  ^ f(i-454) ^ f(i-227)                               ^ f(i)
            //   i = 623        : new[396] ^ g(o[623], new[0])   (both expanded the same way)
            // with f(j) = g(o[j], o[j+1]); the state is double buffered, so all 624 words are computed in one phase.
            const uint32_t* o = mt_buf[cur];
            uint32_t* w = mt_buf[cur ^ 1];
            auto f = [&](int q) { return mt_mix(o[q], o[q + 1], 0u); };
            for (int i = tid; i < kMtN; i += kStreamThreads) {
                uint32_t v;
                if (i < kD) v = o[i + kMtM] ^ f(i);
                else if (i < 2 * kD) v = o[i + kMtM - kD] ^ f(i - kD) ^ f(i);
                else if (i < kMtN - 1) v = o[i + kMtM - 2 * kD] ^ f(i - 2 * kD) ^ f(i - kD) ^ f(i);
                else {
                    const uint32_t n0 = o[kMtM] ^ f(0);                                              // new[0]
                    const uint32_t n396 = o[kMtM - 1 + kMtM - kD] ^ f(kMtM - 1 - kD) ^ f(kMtM - 1);  // new[396]
                    v = mt_mix(o[kMtN - 1], n0, n396);
                }
                w[i] = v;
            }
            __syncthreads();
            cur ^= 1;
            pos = 0;
        }
        const uint32_t* mt = mt_buf[cur];
        const unsigned int blk_end = min(consumed + (unsigned int)(kMtN - pos), T);
        while (lo_j < blk_end) {  // (hi_j > consumed always holds for the current job)
            const unsigned int lo = max(lo_j, consumed), hi = min(hi_j, blk_end);
            for (unsigned int s = lo + tid; s < hi; s += kStreamThreads)
                draws[(size_t)j * S + (s - lo_j)] = mt_temper(mt[pos + (int)(s - consumed)]);
            if (hi_j > blk_end) break;
            ++j;
            next_job();
        }
        pos += (int)(blk_end - consumed);
        consumed = blk_end;
        // no barrier here: the next twist writes the OTHER buffer, whose last readers finished before the barrier above
    }
    __syncthreads();
    for (int i = tid; i < kMtN; i += kStreamThreads) state[i] = mt_buf[cur][i];
    if (tid == 0) state[kMtN] = (uint32_t)pos;
}

struct ApplySmem {
    int hkey[kHashSize];
    int hval[kHashSize];
    int perm[kMaxTake];
    uint32_t draw[kMaxTake];
};

__device__ __forceinline__ int hash_slot(const int* hkey, int key) {
    unsigned int h = ((unsigned int)key * 2654435761u) >> 21;  // 11 bits
    while (true) {
        const int k = hkey[h];
        if (k == key || k == -1) return (int)h;
        h = (h + 1) & (kHashSize - 1);
    }
}

__global__ void __launch_bounds__(kApplyThreads)
    sample_apply_kernel(const int4* __restrict__ jobs, const uint32_t* __restrict__ draws, int S, int N,
                        int8_t* __restrict__ rpn_label8, const int32_t* __restrict__ rpn_pos_list,
                        const int32_t* __restrict__ rpn_neg_list, int32_t* __restrict__ sel, int32_t* __restrict__ sel_n,
                        int sel_stride) {
    __shared__ ApplySmem sm;
    const int j = blockIdx.x, b = j >> 2, kind = j & 3, tid = threadIdx.x;
    const int4 jb = jobs[j];
    const int n = jb.x, need = jb.y, take = jb.z;
    if (kind < 2 && (n == 0 || rpn_label8 == nullptr)) return;  // the reference does not sample here
    if (kind >= 2 && sel == nullptr) return;
    for (int i = tid; i < kHashSize; i += kApplyThreads) sm.hkey[i] = -1;
    // swap partner of step i, i + random() % (n - i): the division is done by all threads, off the serial chain
    for (int i = tid; i < need; i += kApplyThreads) sm.draw[i] = (uint32_t)i + draws[(size_t)j * S + i] % (uint32_t)(n - i);
    __syncthreads();
    const int32_t* list = kind < 2 ? (kind == 0 ? rpn_pos_list : rpn_neg_list) + (size_t)b * N : nullptr;
    int8_t* lb = kind < 2 ? rpn_label8 + (size_t)b * N : nullptr;
    if (tid == 0) {
        // the serial part: `need` Fisher-Yates steps (shared-memory hash map of the displaced entries)
        for (int i = 0; i < need; ++i) {
            const int jj = (int)sm.draw[i];
            const int si = hash_slot(sm.hkey, i);
            const int vi = (sm.hkey[si] == i) ? sm.hval[si] : i;
            const int sj = hash_slot(sm.hkey, jj);
            sm.perm[i] = (sm.hkey[sj] == jj) ? sm.hval[sj] : jj;
            sm.hkey[sj] = jj;
            sm.hval[sj] = vi;
        }
        for (int i = need; i < take; ++i) {  // take == n: the last entry is whatever is left
            const int si = hash_slot(sm.hkey, i);
            sm.perm[i] = (sm.hkey[si] == i) ? sm.hval[si] : i;
        }
    } else if (kind < 2 && tid >= 32) {
        // meanwhile the other warps mark the whole candidate list as ignore (models/model.py:228-236: everything after
        // the first `take` entries of the permutation becomes -1; the kept ones are restored below)
        for (int p = tid - 32; p < n; p += kApplyThreads - 32) lb[list[p]] = -1;
    }
    __syncthreads();
    if (kind < 2) {
        const int8_t v = kind == 0 ? 1 : 0;
        for (int i = tid; i < take; i += kApplyThreads) lb[list[sm.perm[i]]] = v;
    } else {
        const int n_pos = kind == 2 ? 0 : jobs[j - 1].z;
        int32_t* row = sel + (size_t)b * sel_stride;
        for (int i = tid; i < take; i += kApplyThreads) row[n_pos + i] = sm.perm[i];
        if (tid == 0) sel_n[2 * b + (kind - 2)] = n_pos + take;
    }
}

}  // namespace frr

extern "C" int frr_sample_targets(const int32_t* rpn_counts, const int32_t* frcnn_counts, int B, int rpn_batch,
                                  int rpn_max_pos, int frcnn_batch, int frcnn_max_pos, uint32_t* mt_state, int N,
                                  int8_t* rpn_label8, const int32_t* rpn_pos_list, const int32_t* rpn_neg_list,
                                  int32_t* sel, int32_t* sel_n, int sel_stride, int32_t* jobs, uint32_t* draws, int draws_stride,
                                  frr_stream_t stream) {
    using namespace frr;
    FRR_CHECK_ARG(B >= 0 && mt_state && jobs && draws, "frr_sample_targets: null state / workspace");
    FRR_CHECK_ARG(rpn_counts || frcnn_counts, "frr_sample_targets: nothing to sample");
    FRR_CHECK_ARG(!rpn_counts || (rpn_label8 && rpn_pos_list && rpn_neg_list && N > 0), "frr_sample_targets: RPN lists missing");
    FRR_CHECK_ARG(!frcnn_counts || (sel && sel_n && sel_stride >= frcnn_batch), "frr_sample_targets: sel [B,%d] too small", frcnn_batch);
    FRR_CHECK_ARG(rpn_batch <= kMaxTake && frcnn_batch <= kMaxTake && rpn_max_pos <= rpn_batch && frcnn_max_pos <= frcnn_batch &&
                      rpn_max_pos >= 0 && frcnn_max_pos >= 0,
                  "frr_sample_targets: sample sizes must be <= %d", kMaxTake);
    FRR_CHECK_ARG(draws_stride >= (rpn_counts ? rpn_batch : 0) && draws_stride >= (frcnn_counts ? frcnn_batch : 0),
                  "frr_sample_targets: draws_stride %d too small", draws_stride);
    FRR_CHECK_ARG(aligned16(jobs), "frr_sample_targets: jobs must be 16-byte aligned");
    if (B == 0) return FRR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // larger batches go through in slices of 256 images: the launches are stream-ordered, so the generator advances
    // through the images in order exactly as one launch would
    for (int b0 = 0; b0 < B; b0 += kJobsSmem / 4) {
        const int nb = (B - b0 < kJobsSmem / 4) ? B - b0 : kJobsSmem / 4;
        int4* jb = (int4*)jobs + 4 * (size_t)b0;
        uint32_t* dr = draws + 4 * (size_t)b0 * draws_stride;
        sample_stream_kernel<<<1, kStreamThreads, 0, st>>>(rpn_counts ? rpn_counts + 2 * (size_t)b0 : nullptr,
                                                           frcnn_counts ? frcnn_counts + 2 * (size_t)b0 : nullptr, nb, rpn_batch,
                                                           rpn_max_pos, frcnn_batch, frcnn_max_pos, mt_state, jb, dr, draws_stride);
        count_launch();
        FRR_CHECK_LAUNCH("sample_stream_kernel");
        sample_apply_kernel<<<4 * nb, kApplyThreads, 0, st>>>(
            jb, dr, draws_stride, N, rpn_label8 ? rpn_label8 + (size_t)b0 * N : nullptr,
            rpn_pos_list ? rpn_pos_list + (size_t)b0 * N : nullptr, rpn_neg_list ? rpn_neg_list + (size_t)b0 * N : nullptr,
            sel ? sel + (size_t)b0 * sel_stride : nullptr, sel_n ? sel_n + 2 * (size_t)b0 : nullptr, sel_stride);
        count_launch();
        FRR_CHECK_LAUNCH("sample_apply_kernel");
    }
    return FRR_OK;
}
