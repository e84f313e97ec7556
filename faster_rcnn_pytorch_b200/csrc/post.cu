// D1/D2: detection post-processing of FRCNN.predict (models/model.py:369-402).
//
//   frr_decode_classwise : softmax over the class logits, reg * (0.1,0.1,0.2,0.2), per-class decode
//                          against the roi, corner form, clamp to [0,1]   (:369-378)
//   frr_class_nms        : FRCNN._suppress (:382-402) for a whole batch in three launches and no host
//                          sync (the reference does 2 D2H copies per class):
//       1. one CTA per (image, class): select prob > thres, stable descending sort (bitonic on
//          (~score bits, row) keys in shared memory), gather the class boxes;
//       2. the cluster keep-list NMS kernel of nms.cu over the B*(C-1) sorted lists at IoU 0.3
//          (double-threshold compare: IoU == 0.3f is suppressed, as torchvision's CPU kernel);
//       3. one CTA per image: class-major compaction into [B, cap, 4] boxes / labels (l-1) / scores.
#include "frr_common.cuh"

namespace frr {

int nms_launch(const float* boxes, const int32_t* counts, int B, int n, double iou_thr, int max_keep, int32_t* keep,
               int32_t* keep_count, float* out_boxes, int cluster_size, int threads, long long* dbg, int unit_boxes,
               frr_stream_t stream, const int32_t* gather_idx = nullptr, int src_n = 0);

// ------------------------------------------------------------------------------------------------ D1
// one warp per roi row
__global__ void __launch_bounds__(256)
    decode_classwise_kernel(const float* __restrict__ cls, const float4* __restrict__ reg, const float4* __restrict__ rois,
                            int rows, int C, float4 stdv, float* __restrict__ prob, float4* __restrict__ boxes) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = cls + (size_t)row * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = __fadd_rn(s, expf(__fsub_rn(x[c], m)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    const float4 r = rois[row];
    // xy_to_cxcy(roi)
    const float rcx = __fmul_rn(__fadd_rn(r.z, r.x), 0.5f), rcy = __fmul_rn(__fadd_rn(r.w, r.y), 0.5f);
    const float rw = __fsub_rn(r.z, r.x), rh = __fsub_rn(r.w, r.y);
    for (int c = lane; c < C; c += 32) {
        prob[(size_t)row * C + c] = __fdiv_rn(expf(__fsub_rn(x[c], m)), s);
        float4 t = reg[(size_t)row * C + c];
        t.x = __fmul_rn(t.x, stdv.x); t.y = __fmul_rn(t.y, stdv.y); t.z = __fmul_rn(t.z, stdv.z); t.w = __fmul_rn(t.w, stdv.w);
        const float cx = __fadd_rn(__fmul_rn(t.x, rw), rcx), cy = __fadd_rn(__fmul_rn(t.y, rh), rcy);
        const float w = __fmul_rn(expf(t.z), rw), h = __fmul_rn(expf(t.w), rh);
        const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
        float4 b;
        b.x = fminf(fmaxf(__fsub_rn(cx, hw), 0.f), 1.f);
        b.y = fminf(fmaxf(__fsub_rn(cy, hh), 0.f), 1.f);
        b.z = fminf(fmaxf(__fadd_rn(cx, hw), 0.f), 1.f);
        b.w = fminf(fmaxf(__fadd_rn(cy, hh), 0.f), 1.f);
        boxes[(size_t)row * C + c] = b;
    }
}

// ------------------------------------------------------------------------------------------------ D2
// problem p = image * (C-1) + (l-1).  prob [B,R,C], boxes [B,R,C,4].  Output lists have stride R.
__global__ void __launch_bounds__(256)
    class_sort_kernel(const float* __restrict__ prob, const float4* __restrict__ boxes, const int32_t* __restrict__ roi_count,
                      int R, int C, int P2 /* pow2 >= R */, float thres, float4* __restrict__ sboxes,
                      float* __restrict__ sscores, int32_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ int s_n;
    const int p = blockIdx.x;
    const int img = p / (C - 1), l = p % (C - 1) + 1;
    const int Rn = roi_count ? min(roi_count[img], R) : R;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int mine = 0;
    for (int r = threadIdx.x; r < P2; r += blockDim.x) {
        unsigned long long k = ~0ull;
        if (r < Rn) {
            const float s = prob[((size_t)img * R + r) * C + l];
            if (s > thres) {  // fp32 compare (:388)
                k = ((unsigned long long)(~float_to_ordered(s)) << 32) | (unsigned int)r;
                ++mine;
            }
        }
        keys[r] = k;
    }
    if (mine) atomicAdd(&s_n, mine);
    __syncthreads();
    for (int size = 2; size <= P2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int q = threadIdx.x; q < (P2 >> 1); q += blockDim.x) {
                const int lo = 2 * q - (q & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = keys[lo], c2 = keys[hi];
                if ((a > c2) == up) { keys[lo] = c2; keys[hi] = a; }
            }
            __syncthreads();
        }
    const int n = s_n;
    for (int j = threadIdx.x; j < R; j += blockDim.x) {
        const size_t o = (size_t)p * R + j;
        if (j < n) {
            const unsigned long long e = keys[j];
            const int r = (int)(unsigned int)(e & 0xffffffffu);
            sscores[o] = ordered_to_float(~(unsigned int)(e >> 32));
            sboxes[o] = boxes[((size_t)img * R + r) * C + l];
        } else {
            sscores[o] = 0.f;
            sboxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (threadIdx.x == 0) counts[p] = n;
}

__global__ void __launch_bounds__(256)
    det_compact_kernel(const float4* __restrict__ kept_boxes, const float* __restrict__ sscores,
                       const int32_t* __restrict__ keep, const int32_t* __restrict__ keep_count, int R, int C, int cap,
                       float4* __restrict__ det_boxes, int32_t* __restrict__ det_labels, float* __restrict__ det_scores,
                       int32_t* __restrict__ det_count) {
    __shared__ int s_off[1024];
    const int img = blockIdx.x;
    const int nc = C - 1;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int l = 0; l < nc; ++l) {
            s_off[l] = run;
            run += keep_count[img * nc + l];
        }
        s_off[nc] = run;
        det_count[img] = min(run, cap);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int l = warp; l < nc; l += nwarps) {
        const int p = img * nc + l;
        const int n = keep_count[p], off = s_off[l];
        for (int j = lane; j < n; j += 32) {
            const int d = off + j;
            if (d < cap) {
                const size_t o = (size_t)img * cap + d;
                det_boxes[o] = kept_boxes[(size_t)p * R + j];
                det_scores[o] = sscores[(size_t)p * R + keep[(size_t)p * R + j]];
                det_labels[o] = l;  // class l+1 -> label l (:397)
            }
        }
    }
    const int total = min(s_off[nc], cap);
    for (int d = total + threadIdx.x; d < cap; d += blockDim.x) {
        const size_t o = (size_t)img * cap + d;
        det_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        det_scores[o] = 0.f;
        det_labels[o] = -1;
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Evaluation hand-off (test.py:68-88 + evaluation/coco_eval.py:70-92,156-158): scale the normalised boxes to pixels,
// optionally convert xyxy -> xywh, and pack the first max_det detections of every image into fixed-shape rows
// (x, y, x2|w, y2|h, score, label) for ONE tensor all-gather instead of a pickled one.
__global__ void __launch_bounds__(128)
    pack_detections_kernel(const float4* __restrict__ det_boxes, const int32_t* __restrict__ det_labels,
                           const float* __restrict__ det_scores, const int32_t* __restrict__ det_count, int cap,
                           int max_det, const float2* __restrict__ image_wh, int xywh, float* __restrict__ out,
                           int32_t* __restrict__ out_count) {
    const int b = blockIdx.x;
    const int n = min(min(det_count[b], cap), max_det);
    const float2 wh = image_wh ? image_wh[b] : make_float2(1.f, 1.f);
    for (int j = threadIdx.x; j < max_det; j += blockDim.x) {
        float* o = out + ((size_t)b * max_det + j) * 6;
        if (j < n) {
            const float4 bx = det_boxes[(size_t)b * cap + j];
            const float x1 = __fmul_rn(bx.x, wh.x), y1 = __fmul_rn(bx.y, wh.y);
            const float x2 = __fmul_rn(bx.z, wh.x), y2 = __fmul_rn(bx.w, wh.y);
            o[0] = x1;
            o[1] = y1;
            o[2] = xywh ? __fsub_rn(x2, x1) : x2;
            o[3] = xywh ? __fsub_rn(y2, y1) : y2;
            o[4] = det_scores[(size_t)b * cap + j];
            o[5] = (float)det_labels[(size_t)b * cap + j];
        } else {
#pragma unroll
            for (int q = 0; q < 6; ++q) o[q] = 0.f;
        }
    }
    if (threadIdx.x == 0) out_count[b] = n;
}

}  // namespace frr

extern "C" {

int frr_decode_classwise(const float* cls_logits, const float* reg, const float* rois, int rows, int C,
                         const float* std4_host, float* prob, float* boxes, frr_stream_t stream) {
    using namespace frr;
    FRR_CHECK_ARG(rows >= 0 && C >= 1, "frr_decode_classwise: bad sizes");
    if (rows == 0) return FRR_OK;
    FRR_CHECK_ARG(cls_logits && reg && rois && prob && boxes && std4_host, "frr_decode_classwise: null pointer");
    FRR_CHECK_ARG(aligned16(reg) && aligned16(rois) && aligned16(boxes), "frr_decode_classwise: box arrays must be 16-byte aligned");
    const float4 stdv = make_float4(std4_host[0], std4_host[1], std4_host[2], std4_host[3]);
    decode_classwise_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(cls_logits, (const float4*)reg,
                                                                               (const float4*)rois, rows, C, stdv, prob,
                                                                               (float4*)boxes);
    count_launch();
    FRR_CHECK_LAUNCH("decode_classwise_kernel");
    return FRR_OK;
}

int frr_pack_detections(const float* det_boxes, const int32_t* det_labels, const float* det_scores, const int32_t* det_count,
                        int B, int cap, int max_det, const float* image_wh, int xywh, float* out, int32_t* out_count,
                        frr_stream_t stream) {
    using namespace frr;
    FRR_CHECK_ARG(B >= 0 && cap >= 0 && max_det >= 1, "frr_pack_detections: bad sizes");
    if (B == 0) return FRR_OK;
    FRR_CHECK_ARG(det_boxes && det_labels && det_scores && det_count && out && out_count, "frr_pack_detections: null pointer");
    FRR_CHECK_ARG(aligned16(det_boxes) && (image_wh == nullptr || (reinterpret_cast<uintptr_t>(image_wh) & 7u) == 0),
                  "frr_pack_detections: det_boxes must be 16-byte aligned, image_wh 8-byte aligned");
    pack_detections_kernel<<<B, 128, 0, (cudaStream_t)stream>>>((const float4*)det_boxes, det_labels, det_scores, det_count,
                                                                cap, max_det, (const float2*)image_wh, xywh, out, out_count);
    count_launch();
    FRR_CHECK_LAUNCH("pack_detections_kernel");
    return FRR_OK;
}

size_t frr_class_nms_workspace_bytes(int B, int R, int C) {
    using namespace frr;
    const size_t P = (size_t)B * (size_t)(C > 1 ? C - 1 : 0);
    return align256(P * R * 16) * 2 + align256(P * R * 4) * 2 + align256(P * 4) * 2 + 256;
}

int frr_class_nms(const float* prob, const float* boxes, const int32_t* roi_count, int B, int R, int C, float score_thres,
                  double iou_thr, int cap, float* det_boxes, int32_t* det_labels, float* det_scores, int32_t* det_count,
                  void* workspace, size_t workspace_bytes, frr_stream_t stream) {
    using namespace frr;
    FRR_CHECK_ARG(B >= 0 && R >= 0 && C >= 2 && cap >= 0, "frr_class_nms: bad sizes");
    FRR_CHECK_ARG(R <= 4096 && C <= 1024, "frr_class_nms: R <= 4096 and C <= 1024 supported");
    if (B == 0) return FRR_OK;
    FRR_CHECK_ARG(prob && boxes && det_boxes && det_labels && det_scores && det_count && workspace, "frr_class_nms: null pointer");
    FRR_CHECK_ARG(aligned16(boxes) && aligned16(det_boxes) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
                  "frr_class_nms: boxes must be 16-byte aligned, workspace 256-byte aligned");
    if (workspace_bytes < frr_class_nms_workspace_bytes(B, R, C)) {
        set_error("frr_class_nms: workspace too small (%zu < %zu)", workspace_bytes, frr_class_nms_workspace_bytes(B, R, C));
        return FRR_E_WORKSPACE;
    }
    const size_t P = (size_t)B * (C - 1);
    char* w = (char*)workspace;
    float4* sboxes = (float4*)w;          w += align256(P * R * 16);
    float4* kboxes = (float4*)w;          w += align256(P * R * 16);
    float* sscores = (float*)w;           w += align256(P * R * 4);
    int32_t* keep = (int32_t*)w;          w += align256(P * R * 4);
    int32_t* counts = (int32_t*)w;        w += align256(P * 4);
    int32_t* keep_count = (int32_t*)w;
    cudaStream_t st = (cudaStream_t)stream;
    if (R > 0) {
        int P2 = 32;
        while (P2 < R) P2 <<= 1;
        class_sort_kernel<<<(unsigned)P, 256, (size_t)P2 * 8, st>>>(prob, (const float4*)boxes, roi_count, R, C, P2,
                                                                    score_thres, sboxes, sscores, counts);
        count_launch();
        FRR_CHECK_LAUNCH("class_sort_kernel");
        int rc = nms_launch((const float*)sboxes, counts, (int)P, R, iou_thr, R, keep, keep_count, (float*)kboxes, 1, 256,
                            nullptr, 0, stream);
        if (rc) return rc;
    } else {
        FRR_CUDA(cudaMemsetAsync(keep_count, 0, P * 4, st));
    }
    det_compact_kernel<<<B, 256, 0, st>>>(kboxes, sscores, keep, keep_count, R, C, cap, (float4*)det_boxes, det_labels,
                                          det_scores, det_count);
    count_launch();
    FRR_CHECK_LAUNCH("det_compact_kernel");
    return FRR_OK;
}

}  // extern "C"
