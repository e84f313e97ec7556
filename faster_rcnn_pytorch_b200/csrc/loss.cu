// Loss side of the region stage: losses/loss.py:5-59 (SmoothL1Loss, RPNLoss, FastRCNNLoss) together
// with the class-row gather of models/model.py:340-341, consuming the target tensors of the target makers directly.
//
// One CTA per image, one pass: the four losses AND their gradients w.r.t. the four prediction tensors (for a unit
// upstream gradient; the autograd wrapper scales them) are produced together -- the predictions are read once, the
// gradients written once, the [S,C,4] -> [S,4] class-row gather is never materialised.
//   rpn_cls  = CrossEntropy(ignore_index = -1)(cls [N,2], t [N])                 mean over t >= 0      (loss.py:32)
//   rpn_reg  = sum SmoothL1(beta = 1/9)(reg[t > 0] - tg[t > 0]) / #(t >= 0)                             (loss.py:33-38)
//   frc_cls  = CrossEntropy(cls [S,C], c [S])                                     mean over the samples (loss.py:55)
//   frc_reg  = sum SmoothL1(beta = 1)(reg[s, c_s][c > 0] - tg[c > 0]) / #(c >= 0)                       (loss.py:56-59)
// Rows with a negative class in the Fast R-CNN targets are padding of a short sample (fewer than S rois) and are skipped.
// Reductions are fixed-order (per-thread partial sums, warp shuffles, one smem pass): deterministic run to run.
#include <float.h>

#include "frr_common.cuh"

namespace frr {

constexpr int kLossThreads = 1024;

__device__ __forceinline__ float smooth_l1(float d, float beta, float* grad) {
    const float x = fabsf(d);
    if (x >= beta) {  // losses/loss.py:11-14: where(x >= beta, x - 0.5 beta, 0.5 x^2 / beta)
        *grad = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        return x - 0.5f * beta;
    }
    *grad = d / beta;
    return 0.5f * x * x / beta;
}

// deterministic block sum of up to 4 values per thread; result valid in every thread
__device__ __forceinline__ void block_sum4(float v[4], float (*tmp)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < 4; ++q) tmp[warp][q] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float s = 0.f;
        for (int w = 0; w < kLossThreads / 32; ++w) s += tmp[w][q];
        v[q] = s;
    }
}

__global__ void __launch_bounds__(kLossThreads)
    region_loss_kernel(const float2* __restrict__ rpn_cls, const float4* __restrict__ rpn_reg,
                       const int64_t* __restrict__ rpn_tcls, const float4* __restrict__ rpn_treg, int N,
                       const float* __restrict__ frc_cls, const float4* __restrict__ frc_reg,
                       const int64_t* __restrict__ frc_tcls, const float4* __restrict__ frc_treg, int S, int C,
                       int CR /* class rows of frc_reg per sample: C, or 1 when already gathered */, float rpn_beta, float frc_beta, float* __restrict__ loss /* [B,5] */,
                       float2* __restrict__ g_rpn_cls, float4* __restrict__ g_rpn_reg, float* __restrict__ g_frc_cls,
                       float4* __restrict__ g_frc_reg) {
    __shared__ float tmp[kLossThreads / 32][4];
    const int b = blockIdx.x, tid = threadIdx.x;
    float out[5] = {0.f, 0.f, 0.f, 0.f, 0.f};

    if (rpn_cls) {
        const int64_t* t = rpn_tcls + (size_t)b * N;
        // pass 1: #(t >= 0) (the normaliser of both RPN losses)
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < N; i += kLossThreads) acc[0] += (t[i] >= 0) ? 1.f : 0.f;
        block_sum4(acc, tmp);
        const float nv = acc[0];
        const float inv = 1.0f / nv;  // inf when nothing is sampled: the losses become NaN like the reference's 0 / 0
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        for (int i = tid; i < N; i += kLossThreads) {
            const size_t o = (size_t)b * N + i;
            const int64_t ti = t[i];
            float2 gc = make_float2(0.f, 0.f);
            float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ti >= 0) {
                const float2 l = rpn_cls[o];
                const float m = fmaxf(l.x, l.y);
                const float e0 = expf(l.x - m), e1 = expf(l.y - m);
                const float lse = m + logf(e0 + e1);
                acc[0] += lse - (ti == 0 ? l.x : l.y);
                const float is = 1.0f / (e0 + e1);
                gc.x = (e0 * is - (ti == 0 ? 1.f : 0.f)) * inv;
                gc.y = (e1 * is - (ti == 0 ? 0.f : 1.f)) * inv;
                if (ti > 0) {
                    const float4 p = rpn_reg[o], q = rpn_treg[o];
                    float g;
                    acc[1] += smooth_l1(p.x - q.x, rpn_beta, &g); gr.x = g * inv;
                    acc[1] += smooth_l1(p.y - q.y, rpn_beta, &g); gr.y = g * inv;
                    acc[1] += smooth_l1(p.z - q.z, rpn_beta, &g); gr.z = g * inv;
                    acc[1] += smooth_l1(p.w - q.w, rpn_beta, &g); gr.w = g * inv;
                }
            }
            if (g_rpn_cls) g_rpn_cls[o] = gc;
            if (g_rpn_reg) g_rpn_reg[o] = gr;
        }
        block_sum4(acc, tmp);
        out[1] = acc[0] / nv;
        out[2] = acc[1] / nv;
    }

    if (frc_cls) {
        const int64_t* t = frc_tcls + (size_t)b * S;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = tid; s < S; s += kLossThreads) acc[0] += (t[s] >= 0) ? 1.f : 0.f;
        block_sum4(acc, tmp);
        const float nv = acc[0];
        const float inv = 1.0f / nv;
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        // one warp per sample row: C logits over the lanes
        const int lane = tid & 31, warp = tid >> 5;
        for (int s = warp; s < S; s += kLossThreads / 32) {
            const size_t row = (size_t)b * S + s;
            const int64_t ts = t[s];
            const float* l = frc_cls + row * C;
            float m = -FLT_MAX;
            for (int c = lane; c < C; c += 32) m = fmaxf(m, l[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float se = 0.f;
            for (int c = lane; c < C; c += 32) se += expf(l[c] - m);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
            const bool use = ts >= 0 && ts < C;
            if (use && lane == 0) acc[0] += m + logf(se) - l[ts];
            if (g_frc_cls) {
                const float is = use ? inv / se : 0.f;
                for (int c = lane; c < C; c += 32)
                    g_frc_cls[row * C + c] = use ? (expf(l[c] - m) * is - (c == ts ? inv : 0.f)) : 0.f;
            }
            // class-row gather (models/model.py:340-341) + SmoothL1 on the positives
            float4 gsel = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use && ts > 0 && lane == 0) {
                const float4 p = frc_reg[row * CR + (CR == 1 ? 0 : ts)], q = frc_treg[row];
                float g;
                acc[1] += smooth_l1(p.x - q.x, frc_beta, &g); gsel.x = g * inv;
                acc[1] += smooth_l1(p.y - q.y, frc_beta, &g); gsel.y = g * inv;
                acc[1] += smooth_l1(p.z - q.z, frc_beta, &g); gsel.z = g * inv;
                acc[1] += smooth_l1(p.w - q.w, frc_beta, &g); gsel.w = g * inv;
            }
            if (g_frc_reg) {
                gsel.x = __shfl_sync(0xffffffffu, gsel.x, 0); gsel.y = __shfl_sync(0xffffffffu, gsel.y, 0);
                gsel.z = __shfl_sync(0xffffffffu, gsel.z, 0); gsel.w = __shfl_sync(0xffffffffu, gsel.w, 0);
                for (int c = lane; c < CR; c += 32)
                    g_frc_reg[row * CR + c] = (use && ts > 0 && (CR == 1 || c == ts)) ? gsel : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        block_sum4(acc, tmp);
        out[3] = acc[0] / nv;
        out[4] = acc[1] / nv;
    }
    if (tid == 0) {
        out[0] = ((out[1] + out[2]) + out[3]) + out[4];  // losses/loss.py:81
#pragma unroll
        for (int q = 0; q < 5; ++q) loss[5 * b + q] = out[q];
    }
}

}  // namespace frr

extern "C" int frr_region_loss(const float* rpn_cls, const float* rpn_reg, const int64_t* rpn_target_cls,
                               const float* rpn_target_reg, int B, int N, const float* frcnn_cls, const float* frcnn_reg,
                               const int64_t* frcnn_target_cls, const float* frcnn_target_reg, int S, int C,
                               int frcnn_reg_rows, float rpn_beta, float frcnn_beta, float* loss, float* grad_rpn_cls, float* grad_rpn_reg,
                               float* grad_frcnn_cls, float* grad_frcnn_reg, frr_stream_t stream) {
    using namespace frr;
    FRR_CHECK_ARG(B >= 0 && loss, "frr_region_loss: null loss / bad B");
    FRR_CHECK_ARG(rpn_cls || frcnn_cls, "frr_region_loss: nothing to compute");
    FRR_CHECK_ARG(!rpn_cls || (rpn_reg && rpn_target_cls && rpn_target_reg && N > 0), "frr_region_loss: RPN tensors missing");
    FRR_CHECK_ARG(!frcnn_cls || (frcnn_reg && frcnn_target_cls && frcnn_target_reg && S > 0 && C > 0),
                  "frr_region_loss: Fast R-CNN tensors missing");
    FRR_CHECK_ARG(!frcnn_cls || frcnn_reg_rows == C || frcnn_reg_rows == 1, "frr_region_loss: frcnn_reg_rows must be C or 1");
    FRR_CHECK_ARG(rpn_beta > 0.f && frcnn_beta > 0.f, "frr_region_loss: beta must be positive");
    FRR_CHECK_ARG(aligned16(rpn_reg) && aligned16(rpn_target_reg) && aligned16(frcnn_reg) && aligned16(frcnn_target_reg) &&
                      aligned16(grad_rpn_reg) && aligned16(grad_frcnn_reg),
                  "frr_region_loss: box tensors must be 16-byte aligned");
    if (B == 0) return FRR_OK;
    region_loss_kernel<<<B, kLossThreads, 0, (cudaStream_t)stream>>>(
        (const float2*)rpn_cls, (const float4*)rpn_reg, rpn_target_cls, (const float4*)rpn_target_reg, N, frcnn_cls,
        (const float4*)frcnn_reg, frcnn_target_cls, (const float4*)frcnn_target_reg, S, C, frcnn_reg_rows, rpn_beta, frcnn_beta, loss,
        (float2*)grad_rpn_cls, (float4*)grad_rpn_reg, grad_frcnn_cls, (float4*)grad_frcnn_reg);
    count_launch();
    FRR_CHECK_LAUNCH("region_loss_kernel");
    return FRR_OK;
}
