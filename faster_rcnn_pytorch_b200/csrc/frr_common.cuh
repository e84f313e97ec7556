// Shared helpers for libfrr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "frr.h"

namespace frr {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define FRR_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ::frr::set_error(__VA_ARGS__);  \
            return FRR_E_INVALID;           \
        }                                   \
    } while (0)

#define FRR_CHECK_LAUNCH(what)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            ::frr::set_error("%s: %s", what, cudaGetErrorString(e__));                  \
            return FRR_E_LAUNCH;                                                        \
        }                                                                               \
    } while (0)

#define FRR_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ::frr::set_error("%s: %s", #call, cudaGetErrorString(e__));                 \
            return FRR_E_LAUNCH;                                                        \
        }                                                                               \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// streaming 128-bit load / store (read-once inputs, write-once outputs)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// fp32 <-> order-preserving uint32 (ascending uint == ascending float; -0 < +0)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct AnchorTable {
    float v[16 * 4];
    int A;
};

int fill_anchor_table(AnchorTable* t, const float* base_table_host, int A, int stride);

}  // namespace frr
