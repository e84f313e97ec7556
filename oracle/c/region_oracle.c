/*
 * CPU oracle, C part  --  TEST INFRASTRUCTURE ONLY (see oracle/region_oracle.py header).
 *
 * Plain-C restatement of the third-party CPU kernels the reference calls on its region
 * stage.  The reference (csm-kr/faster_rcnn_pytorch) has no native code; the arithmetic
 * lives in torchvision (unpinned by the reference; 0.26.0 in the build container):
 *   - torchvision::nms            called at models/model.py:53 and :394
 *   - torchvision::roi_pool       called at models/model.py:113 (RoIPool((7,7),1.0) :97)
 *   - torchvision::roi_align      called via MultiScaleRoIAlign at models/new_model.py:127,143
 * Semantics restated from the published algorithm and
 * pinned bit-exact against torchvision's CPU kernels by tests/golden (make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * No FMA contraction, IEEE division: results must not depend on the compiler.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ NMS (row N2) ------ */
/* boxes [n,4] fp32 in ORIGINAL order; order[n] = indices sorted by descending score
 * (stable).  Writes kept original indices to keep_out (capacity n), returns the count.
 * Suppress j when (double)(inter / ((area_i + area_j) - inter)) > thr.               */
int64_t oracle_nms_sorted(const float* boxes, const int64_t* order, int64_t n, double thr,
                          int64_t* keep_out)
{
    if (n <= 0) return 0;
    uint8_t* dead = (uint8_t*)calloc((size_t)n, 1);
    float* area = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    int64_t nk = 0;
    for (int64_t oi = 0; oi < n; ++oi) {
        const int64_t i = order[oi];
        if (dead[i]) continue;
        keep_out[nk++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1];
        const float ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float ia = area[i];
        for (int64_t oj = oi + 1; oj < n; ++oj) {
            const int64_t j = order[oj];
            if (dead[j]) continue;
            const float xx1 = fmaxf(ix1, boxes[4 * j]);
            const float yy1 = fmaxf(iy1, boxes[4 * j + 1]);
            const float xx2 = fminf(ix2, boxes[4 * j + 2]);
            const float yy2 = fminf(iy2, boxes[4 * j + 3]);
            const float w = fmaxf(0.0f, xx2 - xx1);
            const float h = fmaxf(0.0f, yy2 - yy1);
            const float inter = w * h;
            const float ovr = inter / ((ia + area[j]) - inter);
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    free(dead);
    free(area);
    return nk;
}

/* ------------------------------------------------------------- RoIPool (rows R2/R3) --- */
/* feat [B,C,H,W], rois [K,5] = (batch, x1,y1,x2,y2); out/argmax [K,C,PH,PW].             */
void oracle_roi_pool_fwd(const float* feat, const float* rois, int64_t K, int64_t C, int64_t H,
                         int64_t W, int PH, int PW, float scale, float* out, int32_t* argmax)
{
    for (int64_t k = 0; k < K; ++k) {
        const float* r = rois + 5 * k;
        const int64_t b = (int64_t)r[0];
        const int sw = (int)roundf(r[1] * scale), sh = (int)roundf(r[2] * scale);
        const int ew = (int)roundf(r[3] * scale), eh = (int)roundf(r[4] * scale);
        const int rw = (ew - sw + 1) > 1 ? (ew - sw + 1) : 1;
        const int rh = (eh - sh + 1) > 1 ? (eh - sh + 1) : 1;
        const float bin_h = (float)rh / (float)PH;
        const float bin_w = (float)rw / (float)PW;
        for (int ph = 0; ph < PH; ++ph) {
            for (int pw = 0; pw < PW; ++pw) {
                int hs = (int)floorf((float)ph * bin_h);
                int ws = (int)floorf((float)pw * bin_w);
                int he = (int)ceilf((float)(ph + 1) * bin_h);
                int we = (int)ceilf((float)(pw + 1) * bin_w);
                hs = hs + sh; he = he + sh; ws = ws + sw; we = we + sw;
                if (hs < 0) hs = 0; if (hs > H) hs = (int)H;
                if (he < 0) he = 0; if (he > H) he = (int)H;
                if (ws < 0) ws = 0; if (ws > W) ws = (int)W;
                if (we < 0) we = 0; if (we > W) we = (int)W;
                const int empty = (he <= hs) || (we <= ws);
                for (int64_t c = 0; c < C; ++c) {
                    const float* plane = feat + ((b * C + c) * H) * W;
                    float best = empty ? 0.0f : -FLT_MAX;
                    int32_t bi = -1;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) {
                            const float v = plane[(int64_t)h * W + w];
                            if (v > best) { best = v; bi = (int32_t)(h * W + w); }
                        }
                    const int64_t o = ((k * C + c) * PH + ph) * PW + pw;
                    out[o] = best;
                    argmax[o] = bi;
                }
            }
        }
    }
}

void oracle_roi_pool_bwd(const float* grad_out, const int32_t* argmax, const float* rois, int64_t K,
                         int64_t C, int64_t H, int64_t W, int PH, int PW, float* grad_in /* zeroed */)
{
    for (int64_t k = 0; k < K; ++k) {
        const int64_t b = (int64_t)rois[5 * k];
        for (int64_t c = 0; c < C; ++c) {
            float* plane = grad_in + ((b * C + c) * H) * W;
            for (int i = 0; i < PH * PW; ++i) {
                const int64_t o = (k * C + c) * PH * PW + i;
                if (argmax[o] != -1) plane[argmax[o]] += grad_out[o];
            }
        }
    }
}

/* ------------------------------------------------------------------ RoIAlign (row R4) - */
typedef struct { int p1, p2, p3, p4; float w1, w2, w3, w4; } tap_t;

static void bilinear_taps(float y, float x, int H, int W, tap_t* t)
{
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
        t->p1 = t->p2 = t->p3 = t->p4 = -1;
        t->w1 = t->w2 = t->w3 = t->w4 = 0.0f;
        return;
    }
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
    const float ly = y - (float)yl, lx = x - (float)xl;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    t->w1 = hy * hx; t->w2 = hy * lx; t->w3 = ly * hx; t->w4 = ly * lx;
    t->p1 = yl * W + xl; t->p2 = yl * W + xh; t->p3 = yh * W + xl; t->p4 = yh * W + xh;
}

typedef struct { float sw, sh, bin_h, bin_w; int gh, gw; float count; int64_t b; } geom_t;

static void roi_geom(const float* r, float scale, int PH, int PW, int sampling, int aligned, geom_t* g)
{
    const float off = aligned ? 0.5f : 0.0f;
    g->b = (int64_t)r[0];
    g->sw = r[1] * scale - off;
    g->sh = r[2] * scale - off;
    const float ew = r[3] * scale - off, eh = r[4] * scale - off;
    float rw = ew - g->sw, rh = eh - g->sh;
    if (!aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    g->bin_h = rh / (float)PH;
    g->bin_w = rw / (float)PW;
    g->gh = sampling > 0 ? sampling : (int)ceilf(rh / (float)PH);
    g->gw = sampling > 0 ? sampling : (int)ceilf(rw / (float)PW);
    const int cnt = g->gh * g->gw;
    g->count = (float)(cnt > 1 ? cnt : 1);
}

void oracle_roi_align_fwd(const float* feat, const float* rois, int64_t K, int64_t C, int64_t H,
                          int64_t W, int PH, int PW, float scale, int sampling, int aligned, float* out)
{
    for (int64_t k = 0; k < K; ++k) {
        geom_t g;
        roi_geom(rois + 5 * k, scale, PH, PW, sampling, aligned, &g);
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw) {
                const int nt = g.gh * g.gw;
                tap_t* taps = (tap_t*)malloc(sizeof(tap_t) * (size_t)(nt > 0 ? nt : 1));
                int q = 0;
                for (int iy = 0; iy < g.gh; ++iy) {
                    const float yy = g.sh + (float)ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.gh;
                    for (int ix = 0; ix < g.gw; ++ix) {
                        const float xx = g.sw + (float)pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.gw;
                        bilinear_taps(yy, xx, (int)H, (int)W, &taps[q++]);
                    }
                }
                for (int64_t c = 0; c < C; ++c) {
                    const float* plane = feat + ((g.b * C + c) * H) * W;
                    float acc = 0.0f;
                    for (int s = 0; s < nt; ++s) {
                        const tap_t* t = &taps[s];
                        if (t->p1 < 0) { acc += 0.0f; continue; }
                        acc += t->w1 * plane[t->p1] + t->w2 * plane[t->p2] + t->w3 * plane[t->p3] +
                               t->w4 * plane[t->p4];
                    }
                    out[((k * C + c) * PH + ph) * PW + pw] = acc / g.count;
                }
                free(taps);
            }
    }
}

void oracle_roi_align_bwd(const float* grad_out, const float* rois, int64_t K, int64_t C, int64_t H,
                          int64_t W, int PH, int PW, float scale, int sampling, int aligned,
                          float* grad_in /* zeroed */)
{
    for (int64_t k = 0; k < K; ++k) {
        geom_t g;
        roi_geom(rois + 5 * k, scale, PH, PW, sampling, aligned, &g);
        for (int64_t c = 0; c < C; ++c) {
            float* plane = grad_in + ((g.b * C + c) * H) * W;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    const float go = grad_out[((k * C + c) * PH + ph) * PW + pw];
                    for (int iy = 0; iy < g.gh; ++iy) {
                        const float yy = g.sh + (float)ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.gh;
                        for (int ix = 0; ix < g.gw; ++ix) {
                            const float xx = g.sw + (float)pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.gw;
                            tap_t t;
                            bilinear_taps(yy, xx, (int)H, (int)W, &t);
                            if (t.p1 < 0) continue;
                            plane[t.p1] += go * t.w1 / g.count;
                            plane[t.p2] += go * t.w2 / g.count;
                            plane[t.p3] += go * t.w3 / g.count;
                            plane[t.p4] += go * t.w4 / g.count;
                        }
                    }
                }
        }
    }
}

/* --------------------------------------------------- batched helpers for the CPU arm --- */
/* Proposal-layer NMS for a batch of images, sequentially (the reference is batch-1; the
 * Python bridge runs slices of the batch on a thread pool, ctypes releases the GIL).  boxes [B,n,4] already sorted by descending score.   */
void oracle_nms_sorted_batch(const float* boxes, int64_t B, int64_t n, double thr, int64_t max_keep,
                             int64_t* keep_out /* [B,max_keep] */, int64_t* count_out /* [B] */)
{
    for (int64_t b = 0; b < B; ++b) {
        int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
        int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
        for (int64_t i = 0; i < n; ++i) order[i] = i;
        int64_t nk = oracle_nms_sorted(boxes + b * n * 4, order, n, thr, keep);
        if (nk > max_keep) nk = max_keep;
        memcpy(keep_out + b * max_keep, keep, sizeof(int64_t) * (size_t)nk);
        count_out[b] = nk;
        free(order);
        free(keep);
    }
}
