// libfrr.so: error plumbing, version, launch counter, anchor base table (host).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "frr_common.cuh"

namespace frr {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// anchor.py:15-32: w = base*scale*sqrt(ratio), h = base*scale*sqrt(1/ratio) in float64, centre
// (base/2, base/2), rows ratio-major over ratios (0.5,1,2) x scales (8,16,32), stored to fp32.
static void reference_base_table(float* t, int base_size) {
    const double ratios[3] = {0.5, 1.0, 2.0};
    const double scales[3] = {8.0, 16.0, 32.0};
    const double c = base_size / 2.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double w = base_size * scales[j] * sqrt(ratios[i]);
            const double h = base_size * scales[j] * sqrt(1.0 / ratios[i]);
            float* r = t + 4 * (i * 3 + j);
            r[0] = (float)(c - w / 2.0);
            r[1] = (float)(c - h / 2.0);
            r[2] = (float)(c + w / 2.0);
            r[3] = (float)(c + h / 2.0);
        }
}

int fill_anchor_table(AnchorTable* t, const float* base_table_host, int A, int stride) {
    if (base_table_host == nullptr) {
        reference_base_table(t->v, stride);
        t->A = 9;
        return FRR_OK;
    }
    FRR_CHECK_ARG(A >= 1 && A <= 16, "anchor table: A=%d out of range [1,16]", A);
    memcpy(t->v, base_table_host, sizeof(float) * 4 * (size_t)A);
    t->A = A;
    return FRR_OK;
}

}  // namespace frr

extern "C" {

int frr_abi_version(void) { return FRR_ABI_VERSION; }
const char* frr_last_error(void) { return frr::t_err; }
uint64_t frr_launch_count(void) { return frr::g_launches.load(); }

// torchvision AnchorGenerator.generate_anchors (TV models/detection/anchor_utils.py; used by models/new_model.py:23-25):
//   h_ratios = sqrt(ratios); w_ratios = 1 / h_ratios; ws = w_ratios * size; hs = h_ratios * size;
//   base = round(stack(-ws, -hs, ws, hs) / 2)     -- all in fp32, torch.round = half to even
int frr_tv_anchor_base_host(float size, const float* aspect_ratios, int n_ratios, float* table_host) {
    FRR_CHECK_ARG(table_host != nullptr && aspect_ratios != nullptr && n_ratios >= 1 && size > 0.f,
                  "frr_tv_anchor_base_host: bad arguments");
    for (int i = 0; i < n_ratios; ++i) {
        volatile float hr = sqrtf(aspect_ratios[i]);   // volatile: keep every intermediate in fp32
        volatile float wr = 1.0f / hr;
        volatile float ws = wr * size, hs = hr * size;
        float* r = table_host + 4 * i;
        r[0] = nearbyintf(-ws / 2.0f);
        r[1] = nearbyintf(-hs / 2.0f);
        r[2] = nearbyintf(ws / 2.0f);
        r[3] = nearbyintf(hs / 2.0f);
    }
    return FRR_OK;
}

int frr_anchor_base_host(float* table_host, int base_size) {
    FRR_CHECK_ARG(table_host != nullptr && base_size > 0, "frr_anchor_base_host: bad arguments");
    frr::reference_base_table(table_host, base_size);
    return FRR_OK;
}

}  // extern "C"
