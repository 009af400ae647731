// Shared helpers for the balf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace balf {

// ---- error reporting: thread-local message, negative = argument error, positive = cudaError_t
int set_error(int code, const char* fmt, ...);
const char* last_error();

#define BALF_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::balf::set_error(-1, __VA_ARGS__);   \
    } while (0)

#define BALF_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return ::balf::set_error((int)e_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                     __FILE__, __LINE__);                                           \
    } while (0)

#define BALF_LAUNCH_OK() BALF_CUDA_OK(cudaGetLastError())

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_launches;
#define BALF_COUNT_LAUNCH(n) (::balf::g_launches += (n))

// per-kernel CUDA-event timing (bench.py's roofline leg): off by default; when on, every launch
// wrapped in a ProfScope is bracketed by two events on the launching stream.
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
extern bool g_prof_on;
struct ProfScope {
    cudaStream_t st;
    bool on;
    ProfScope(const char* name, cudaStream_t s) : st(s), on(g_prof_on) { if (on) prof_begin(name, s); }
    ~ProfScope() { if (on) prof_end(st); }
};

constexpr float kNegInf = -__builtin_huge_valf();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace balf
