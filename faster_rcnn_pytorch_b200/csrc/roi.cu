// R1-R4: RoIPool / RoIAlign 7x7 forward + backward (torchvision semantics; reference call sites
// models/model.py:97,113 and models/new_model.py:127,143).
//
// Design (B200: 227 KB shared memory per CTA, TMA bulk copies):
//   a CTA owns CB whole channel planes of ONE image.  Forward: the planes are brought into shared
//   memory once (NCHW: one contiguous chunk -> cp.async.bulk / TMA with an mbarrier; channels_last:
//   64-byte granules), then every (roi, channel, bin) output of that image is produced by scanning its
//   window out of shared memory and stored straight to the [K,C,7,7] output, where the CB*49 values of
//   a roi are contiguous (coalesced).  Each feature element is read from HBM exactly once and nothing
//   is transposed, so HBM traffic equals the algorithmic bytes (features + out + argmax).
//   Backward: the CTA owns the CB gradient planes in shared memory, accumulates every roi of its
//   image with shared-memory atomics (no global atomics, no memset pass) and writes each plane once.
//   Planes that do not fit shared memory (large FPN levels) take the direct global-memory kernels.
//
// Numerics: RoIPool max/argmax bit-exact (first strict maximum in h-then-w order, -FLT_MAX start,
// empty bin -> 0/-1); RoIAlign sums w1*v1+w2*v2+w3*v3+w4*v4 left to right without FMA contraction.
#include <float.h>

#include "frr_common.cuh"

namespace frr {

constexpr int kRoiThreads = 512;
constexpr int kRoiTile = 256;  // rois staged per pass

// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copy helpers (global -> shared with mbarrier completion; shared -> global)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Load CB planes of image b into smem planes[CB][HW].  NCHW: contiguous chunk (TMA when 16-B aligned),
// NHWC: gather CB consecutive channels per pixel.
__device__ __forceinline__ void load_planes(float* planes, const float* __restrict__ feat, int b, int c0, int CB, int C,
                                            int HW, bool nhwc, uint64_t* bar) {
    const int tid = threadIdx.x;
    if (!nhwc) {
        const float* src = feat + ((size_t)b * C + c0) * HW;
        const size_t bytes = (size_t)CB * HW * sizeof(float);
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((bytes & 15u) == 0);
        if (tma_ok) {
            if (tid == 0) {
                mbar_init(bar, 1);
                mbar_expect_tx(bar, (uint32_t)bytes);
                const uint32_t kMax = 32768;
                for (size_t off = 0; off < bytes; off += kMax) {
                    const uint32_t nb = (uint32_t)((bytes - off) < kMax ? (bytes - off) : kMax);
                    bulk_g2s(reinterpret_cast<char*>(planes) + off, reinterpret_cast<const char*>(src) + off, nb, bar);
                }
            }
            __syncthreads();  // barrier init visible to the waiters
            mbar_wait(bar, 0);
        } else {
            for (int i = tid; i < CB * HW; i += blockDim.x) planes[i] = src[i];
        }
    } else {
        const float* src = feat + (size_t)b * HW * C + c0;
        for (int i = tid; i < CB * HW; i += blockDim.x) {
            const int pix = i / CB, cl = i - pix * CB;
            planes[(size_t)cl * HW + pix] = src[(size_t)pix * C + cl];
        }
    }
    __syncthreads();
}

// Store CB planes from smem to image b (zero-copy layout choice as above).
__device__ __forceinline__ void store_planes(const float* planes, float* __restrict__ out, int b, int c0, int CB, int C,
                                             int HW, bool nhwc) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (!nhwc) {
        float* dst = out + ((size_t)b * C + c0) * HW;
        const size_t bytes = (size_t)CB * HW * sizeof(float);
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
        if (tma_ok) {
            fence_async_smem();  // generic-proxy smem writes -> visible to the async proxy
            __syncthreads();
            if (tid == 0) {
                const uint32_t kMax = 32768;
                for (size_t off = 0; off < bytes; off += kMax) {
                    const uint32_t nb = (uint32_t)((bytes - off) < kMax ? (bytes - off) : kMax);
                    bulk_s2g(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(planes) + off, nb);
                }
                bulk_commit_wait();
            }
        } else {
            for (int i = tid; i < CB * HW; i += blockDim.x) dst[i] = planes[i];
        }
    } else {
        float* dst = out + (size_t)b * HW * C + c0;
        for (int i = tid; i < CB * HW; i += blockDim.x) {
            const int pix = i / CB, cl = i - pix * CB;
            dst[(size_t)pix * C + cl] = planes[(size_t)cl * HW + pix];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
struct PoolGeom {  // torchvision roi_pool: rounded roi, integer bins
    int sw, sh, rw, rh;
};
__device__ __forceinline__ PoolGeom pool_geom(const float* r, float scale) {
    PoolGeom g;
    g.sw = (int)roundf(__fmul_rn(r[1], scale));
    g.sh = (int)roundf(__fmul_rn(r[2], scale));
    const int ew = (int)roundf(__fmul_rn(r[3], scale));
    const int eh = (int)roundf(__fmul_rn(r[4], scale));
    g.rw = max(ew - g.sw + 1, 1);
    g.rh = max(eh - g.sh + 1, 1);
    return g;
}
__device__ __forceinline__ void pool_window(const PoolGeom& g, int ph, int pw, int PH, int PW, int H, int W, int& hs,
                                            int& he, int& ws, int& we) {
    const float bin_h = __fdiv_rn((float)g.rh, (float)PH);
    const float bin_w = __fdiv_rn((float)g.rw, (float)PW);
    hs = (int)floorf(__fmul_rn((float)ph, bin_h)) + g.sh;
    ws = (int)floorf(__fmul_rn((float)pw, bin_w)) + g.sw;
    he = (int)ceilf(__fmul_rn((float)(ph + 1), bin_h)) + g.sh;
    we = (int)ceilf(__fmul_rn((float)(pw + 1), bin_w)) + g.sw;
    hs = min(max(hs, 0), H);
    he = min(max(he, 0), H);
    ws = min(max(ws, 0), W);
    we = min(max(we, 0), W);
}

struct AlignGeom {  // torchvision roi_align
    float sw, sh, bin_h, bin_w, count;
    int gh, gw;
};
__device__ __forceinline__ AlignGeom align_geom(const float* r, float scale, int PH, int PW, int sampling, bool aligned) {
    AlignGeom g;
    const float off = aligned ? 0.5f : 0.0f;
    g.sw = __fsub_rn(__fmul_rn(r[1], scale), off);
    g.sh = __fsub_rn(__fmul_rn(r[2], scale), off);
    const float ew = __fsub_rn(__fmul_rn(r[3], scale), off);
    const float eh = __fsub_rn(__fmul_rn(r[4], scale), off);
    float rw = __fsub_rn(ew, g.sw), rh = __fsub_rn(eh, g.sh);
    if (!aligned) {
        rw = fmaxf(rw, 1.0f);
        rh = fmaxf(rh, 1.0f);
    }
    g.bin_h = __fdiv_rn(rh, (float)PH);
    g.bin_w = __fdiv_rn(rw, (float)PW);
    g.gh = sampling > 0 ? sampling : (int)ceilf(__fdiv_rn(rh, (float)PH));
    g.gw = sampling > 0 ? sampling : (int)ceilf(__fdiv_rn(rw, (float)PW));
    g.count = (float)max(g.gh * g.gw, 1);
    return g;
}
struct Taps {
    int p1, p2, p3, p4;
    float w1, w2, w3, w4;
};
__device__ __forceinline__ bool bilinear_taps(float y, float x, int H, int W, Taps& t) {
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return false;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
    const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
    const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, lx); t.w3 = __fmul_rn(ly, hx); t.w4 = __fmul_rn(ly, lx);
    t.p1 = yl * W + xl; t.p2 = yl * W + xh; t.p3 = yh * W + xl; t.p4 = yh * W + xh;
    return true;
}
__device__ __forceinline__ float sample_y(const AlignGeom& g, int ph, int iy) {
    // roi_start_h + ph*bin_h + (iy + .5f)*bin_h/grid_h   (left to right)
    return __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)ph, g.bin_h)),
                     __fdiv_rn(__fmul_rn((float)iy + .5f, g.bin_h), (float)g.gh));
}
__device__ __forceinline__ float sample_x(const AlignGeom& g, int pw, int ix) {
    return __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)pw, g.bin_w)),
                     __fdiv_rn(__fmul_rn((float)ix + .5f, g.bin_w), (float)g.gw));
}

// Stage the rois of image b found in rois[tile .. tile+kRoiTile) into smem (ids + 5 floats each).
__device__ __forceinline__ int stage_rois(const float* __restrict__ rois, int K, int tile, int b, int* s_id, float* s_roi,
                                          int* s_n) {
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < kRoiTile; t += blockDim.x) {
        const int k = tile + t;
        if (k < K && (int)rois[5 * (size_t)k] == b) {
            const int slot = atomicAdd(s_n, 1);
            s_id[slot] = k;
#pragma unroll
            for (int q = 0; q < 5; ++q) s_roi[5 * slot + q] = rois[5 * (size_t)k + q];
        }
    }
    __syncthreads();
    return *s_n;
}

struct RoiSmemHdr {
    uint64_t bar;
    int n;
    int pad;
    int id[kRoiTile];
    float roi[kRoiTile * 5];
};

// ---------------------------------------------------------------------------------------------
// forward (planes in shared memory)
// ---------------------------------------------------------------------------------------------
template <bool kAlign>
__global__ void __launch_bounds__(kRoiThreads)
    roi_fwd_planes_kernel(const float* __restrict__ feat, const float* __restrict__ rois, int K, int C, int H, int W,
                          int CB, int PH, int PW, float scale, int sampling, int aligned, int nhwc,
                          float* __restrict__ out, int32_t* __restrict__ argmax) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RoiSmemHdr* hd = reinterpret_cast<RoiSmemHdr*>(smem_raw);
    float* planes = reinterpret_cast<float*>(smem_raw + ((sizeof(RoiSmemHdr) + 127) & ~(size_t)127));
    const int b = blockIdx.y, c0 = blockIdx.x * CB;
    const int HW = H * W, bins = PH * PW;
    const int cb = min(CB, C - c0);
    load_planes(planes, feat, b, c0, cb, C, HW, nhwc != 0, &hd->bar);

    for (int tile = 0; tile < K; tile += kRoiTile) {
        const int ns = stage_rois(rois, K, tile, b, hd->id, hd->roi, &hd->n);
        const int items = ns * cb * bins;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int slot = it / (cb * bins);
            const int rem = it - slot * (cb * bins);
            const int cl = rem / bins, bin = rem - cl * bins;
            const int ph = bin / PW, pw = bin - ph * PW;
            const float* pl = planes + (size_t)cl * HW;
            const size_t o = ((size_t)hd->id[slot] * C + c0 + cl) * bins + bin;
            if (!kAlign) {
                const PoolGeom g = pool_geom(hd->roi + 5 * slot, scale);
                int hs, he, ws, we;
                pool_window(g, ph, pw, PH, PW, H, W, hs, he, ws, we);
                const bool empty = (he <= hs) || (we <= ws);
                float best = empty ? 0.f : -FLT_MAX;
                int bi = -1;
                for (int h = hs; h < he; ++h)
                    for (int w = ws; w < we; ++w) {
                        const float v = pl[h * W + w];
                        if (v > best) { best = v; bi = h * W + w; }
                    }
                out[o] = best;
                if (argmax) argmax[o] = bi;
            } else {
                const AlignGeom g = align_geom(hd->roi + 5 * slot, scale, PH, PW, sampling, aligned != 0);
                float acc = 0.f;
                for (int iy = 0; iy < g.gh; ++iy) {
                    const float y = sample_y(g, ph, iy);
                    for (int ix = 0; ix < g.gw; ++ix) {
                        const float x = sample_x(g, pw, ix);
                        Taps t;
                        if (bilinear_taps(y, x, H, W, t)) {
                            const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w1, pl[t.p1]), __fmul_rn(t.w2, pl[t.p2])),
                                                                __fmul_rn(t.w3, pl[t.p3])),
                                                      __fmul_rn(t.w4, pl[t.p4]));
                            acc = __fadd_rn(acc, v);
                        }
                    }
                }
                out[o] = __fdiv_rn(acc, g.count);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward (gradient planes accumulated in shared memory, written once)
// ---------------------------------------------------------------------------------------------
template <bool kAlign>
__global__ void __launch_bounds__(kRoiThreads)
    roi_bwd_planes_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ argmax,
                          const float* __restrict__ rois, int K, int C, int H, int W, int CB, int PH, int PW, float scale,
                          int sampling, int aligned, int nhwc, float* __restrict__ grad_in) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RoiSmemHdr* hd = reinterpret_cast<RoiSmemHdr*>(smem_raw);
    float* planes = reinterpret_cast<float*>(smem_raw + ((sizeof(RoiSmemHdr) + 127) & ~(size_t)127));
    const int b = blockIdx.y, c0 = blockIdx.x * CB;
    const int HW = H * W, bins = PH * PW;
    const int cb = min(CB, C - c0);
    for (int i = threadIdx.x; i < cb * HW; i += blockDim.x) planes[i] = 0.f;
    __syncthreads();

    for (int tile = 0; tile < K; tile += kRoiTile) {
        const int ns = stage_rois(rois, K, tile, b, hd->id, hd->roi, &hd->n);
        const int items = ns * cb * bins;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int slot = it / (cb * bins);
            const int rem = it - slot * (cb * bins);
            const int cl = rem / bins, bin = rem - cl * bins;
            float* pl = planes + (size_t)cl * HW;
            const size_t o = ((size_t)hd->id[slot] * C + c0 + cl) * bins + bin;
            const float go = grad_out[o];
            if (!kAlign) {
                const int a = argmax[o];
                if (a >= 0) atomicAdd(pl + a, go);
            } else {
                const int ph = bin / PW, pw = bin - ph * PW;
                const AlignGeom g = align_geom(hd->roi + 5 * slot, scale, PH, PW, sampling, aligned != 0);
                for (int iy = 0; iy < g.gh; ++iy) {
                    const float y = sample_y(g, ph, iy);
                    for (int ix = 0; ix < g.gw; ++ix) {
                        const float x = sample_x(g, pw, ix);
                        Taps t;
                        if (bilinear_taps(y, x, H, W, t)) {
                            atomicAdd(pl + t.p1, __fdiv_rn(__fmul_rn(go, t.w1), g.count));
                            atomicAdd(pl + t.p2, __fdiv_rn(__fmul_rn(go, t.w2), g.count));
                            atomicAdd(pl + t.p3, __fdiv_rn(__fmul_rn(go, t.w3), g.count));
                            atomicAdd(pl + t.p4, __fdiv_rn(__fmul_rn(go, t.w4), g.count));
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    store_planes(planes, grad_in, b, c0, cb, C, HW, nhwc != 0);
}

// ---------------------------------------------------------------------------------------------
// direct global-memory kernels for planes too large for shared memory (NCHW or NHWC)
// ---------------------------------------------------------------------------------------------
template <bool kAlign>
__global__ void __launch_bounds__(256)
    roi_fwd_direct_kernel(const float* __restrict__ feat, const float* __restrict__ rois, size_t total, int C, int H, int W,
                          int PH, int PW, float scale, int sampling, int aligned, int nhwc, float* __restrict__ out,
                          int32_t* __restrict__ argmax) {
    const int bins = PH * PW;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const int bin = (int)(o % bins);
        const int c = (int)((o / bins) % C);
        const size_t k = o / ((size_t)bins * C);
        const int ph = bin / PW, pw = bin - ph * PW;
        const float* r = rois + 5 * k;
        const int b = (int)r[0];
        const float* base = nhwc ? feat + (size_t)b * H * W * C + c : feat + ((size_t)b * C + c) * H * W;
        const size_t ps = nhwc ? (size_t)C : 1;  // pixel stride
        if (!kAlign) {
            const PoolGeom g = pool_geom(r, scale);
            int hs, he, ws, we;
            pool_window(g, ph, pw, PH, PW, H, W, hs, he, ws, we);
            const bool empty = (he <= hs) || (we <= ws);
            float best = empty ? 0.f : -FLT_MAX;
            int bi = -1;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    const float v = base[(size_t)(h * W + w) * ps];
                    if (v > best) { best = v; bi = h * W + w; }
                }
            out[o] = best;
            if (argmax) argmax[o] = bi;
        } else {
            const AlignGeom g = align_geom(r, scale, PH, PW, sampling, aligned != 0);
            float acc = 0.f;
            for (int iy = 0; iy < g.gh; ++iy) {
                const float y = sample_y(g, ph, iy);
                for (int ix = 0; ix < g.gw; ++ix) {
                    const float x = sample_x(g, pw, ix);
                    Taps t;
                    if (bilinear_taps(y, x, H, W, t)) {
                        const float v = __fadd_rn(
                            __fadd_rn(__fadd_rn(__fmul_rn(t.w1, base[t.p1 * ps]), __fmul_rn(t.w2, base[t.p2 * ps])),
                                      __fmul_rn(t.w3, base[t.p3 * ps])),
                            __fmul_rn(t.w4, base[t.p4 * ps]));
                        acc = __fadd_rn(acc, v);
                    }
                }
            }
            out[o] = __fdiv_rn(acc, g.count);
        }
    }
}

template <bool kAlign>
__global__ void __launch_bounds__(256)
    roi_bwd_direct_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ argmax,
                          const float* __restrict__ rois, size_t total, int C, int H, int W, int PH, int PW, float scale,
                          int sampling, int aligned, int nhwc, float* __restrict__ grad_in /* zeroed */) {
    const int bins = PH * PW;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const int bin = (int)(o % bins);
        const int c = (int)((o / bins) % C);
        const size_t k = o / ((size_t)bins * C);
        const float* r = rois + 5 * k;
        const int b = (int)r[0];
        float* base = nhwc ? grad_in + (size_t)b * H * W * C + c : grad_in + ((size_t)b * C + c) * H * W;
        const size_t ps = nhwc ? (size_t)C : 1;
        const float go = grad_out[o];
        if (!kAlign) {
            const int a = argmax[o];
            if (a >= 0) atomicAdd(base + (size_t)a * ps, go);
        } else {
            const int ph = bin / PW, pw = bin - ph * PW;
            const AlignGeom g = align_geom(r, scale, PH, PW, sampling, aligned != 0);
            for (int iy = 0; iy < g.gh; ++iy) {
                const float y = sample_y(g, ph, iy);
                for (int ix = 0; ix < g.gw; ++ix) {
                    const float x = sample_x(g, pw, ix);
                    Taps t;
                    if (bilinear_taps(y, x, H, W, t)) {
                        atomicAdd(base + t.p1 * ps, __fdiv_rn(__fmul_rn(go, t.w1), g.count));
                        atomicAdd(base + t.p2 * ps, __fdiv_rn(__fmul_rn(go, t.w2), g.count));
                        atomicAdd(base + t.p3 * ps, __fdiv_rn(__fmul_rn(go, t.w3), g.count));
                        atomicAdd(base + t.p4 * ps, __fdiv_rn(__fmul_rn(go, t.w4), g.count));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t roi_smem_bytes(int CB, int HW) {
    return ((sizeof(RoiSmemHdr) + 127) & ~(size_t)127) + (size_t)CB * HW * sizeof(float);
}

// channels per CTA: the largest of {16,8,4,2,1} that fits shared memory while keeping >= one wave of CTAs
static int pick_cb(int B, int C, int HW) {
    const size_t limit = 200 * 1024;
    int best = 0;
    for (int cb = 16; cb >= 1; cb >>= 1) {
        if (roi_smem_bytes(cb, HW) > limit) continue;
        if (best == 0) best = cb;
        const long ctas = (long)B * ((C + cb - 1) / cb);
        if (ctas >= (long)num_sms()) return cb;
        best = cb;
    }
    return best;  // 0 -> does not fit at all
}

template <bool kAlign>
static int roi_forward(const float* feat, const float* rois, int K, int B, int C, int H, int W, int PH, int PW, float scale,
                       int sampling, int aligned, int nhwc, float* out, int32_t* argmax, frr_stream_t stream) {
    FRR_CHECK_ARG(K == 0 || (feat && out && rois), "roi forward: null pointer");
    FRR_CHECK_ARG(K >= 0 && B > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && B <= 65535, "roi forward: bad sizes");
    if (K == 0) return FRR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    const int cb = pick_cb(B, C, HW);
    if (cb > 0) {
        const size_t smem = roi_smem_bytes(cb, HW);
        auto kern = roi_fwd_planes_kernel<kAlign>;
        FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        dim3 grid((C + cb - 1) / cb, B);
        kern<<<grid, kRoiThreads, smem, st>>>(feat, rois, K, C, H, W, cb, PH, PW, scale, sampling, aligned, nhwc, out,
                                              argmax);
    } else {
        const size_t total = (size_t)K * C * PH * PW;
        const int blocks = (int)((total + 255) / 256 < (size_t)num_sms() * 16 ? (total + 255) / 256 : (size_t)num_sms() * 16);
        roi_fwd_direct_kernel<kAlign><<<blocks, 256, 0, st>>>(feat, rois, total, C, H, W, PH, PW, scale, sampling, aligned,
                                                              nhwc, out, argmax);
    }
    count_launch();
    FRR_CHECK_LAUNCH("roi forward kernel");
    return FRR_OK;
}

template <bool kAlign>
static int roi_backward(const float* grad_out, const int32_t* argmax, const float* rois, int K, int B, int C, int H, int W,
                        int PH, int PW, float scale, int sampling, int aligned, int nhwc, float* grad_in,
                        frr_stream_t stream) {
    FRR_CHECK_ARG(grad_in && (K == 0 || (grad_out && rois)), "roi backward: null pointer");
    FRR_CHECK_ARG(kAlign || K == 0 || argmax, "roi_pool backward: argmax is required");
    FRR_CHECK_ARG(K >= 0 && B > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && B <= 65535, "roi backward: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    const int cb = pick_cb(B, C, HW);
    if (cb > 0) {
        const size_t smem = roi_smem_bytes(cb, HW);
        auto kern = roi_bwd_planes_kernel<kAlign>;
        FRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        dim3 grid((C + cb - 1) / cb, B);
        kern<<<grid, kRoiThreads, smem, st>>>(grad_out, argmax, rois, K, C, H, W, cb, PH, PW, scale, sampling, aligned, nhwc,
                                              grad_in);
        count_launch();
    } else {
        FRR_CUDA(cudaMemsetAsync(grad_in, 0, (size_t)B * C * HW * sizeof(float), st));
        if (K > 0) {
            const size_t total = (size_t)K * C * PH * PW;
            const int blocks =
                (int)((total + 255) / 256 < (size_t)num_sms() * 16 ? (total + 255) / 256 : (size_t)num_sms() * 16);
            roi_bwd_direct_kernel<kAlign><<<blocks, 256, 0, st>>>(grad_out, argmax, rois, total, C, H, W, PH, PW, scale,
                                                                  sampling, aligned, nhwc, grad_in);
            count_launch();
        }
    }
    FRR_CHECK_LAUNCH("roi backward kernel");
    return FRR_OK;
}

}  // namespace frr

extern "C" {

int frr_roi_pool_fwd(const float* feat, const float* rois, int K, int B, int C, int H, int W, int PH, int PW,
                     float spatial_scale, int channels_last, float* out, int32_t* argmax, frr_stream_t stream) {
    return frr::roi_forward<false>(feat, rois, K, B, C, H, W, PH, PW, spatial_scale, 0, 0, channels_last, out, argmax, stream);
}
int frr_roi_pool_bwd(const float* grad_out, const int32_t* argmax, const float* rois, int K, int B, int C, int H, int W,
                     int PH, int PW, float spatial_scale, int channels_last, float* grad_in, frr_stream_t stream) {
    return frr::roi_backward<false>(grad_out, argmax, rois, K, B, C, H, W, PH, PW, spatial_scale, 0, 0, channels_last, grad_in,
                                    stream);
}
int frr_roi_align_fwd(const float* feat, const float* rois, int K, int B, int C, int H, int W, int PH, int PW,
                      float spatial_scale, int sampling_ratio, int aligned, int channels_last, float* out,
                      frr_stream_t stream) {
    return frr::roi_forward<true>(feat, rois, K, B, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned, channels_last, out,
                                  nullptr, stream);
}
int frr_roi_align_bwd(const float* grad_out, const float* rois, int K, int B, int C, int H, int W, int PH, int PW,
                      float spatial_scale, int sampling_ratio, int aligned, int channels_last, float* grad_in,
                      frr_stream_t stream) {
    return frr::roi_backward<true>(grad_out, nullptr, rois, K, B, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned,
                                   channels_last, grad_in, stream);
}

}  // extern "C"
