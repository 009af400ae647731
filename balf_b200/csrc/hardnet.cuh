// HardNet weights: packed layouts shared by the fp32 path (hardnet.cu) and the tensor-core path (hardnet_tc.cu).
#pragma once
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

struct HnLayer { int cin, cout, hin, stride, ks; };
static const HnLayer kHn[7] = {{1, 32, 32, 1, 3},  {32, 32, 32, 1, 3},  {32, 64, 32, 2, 3},  {64, 64, 16, 1, 3},
                               {64, 128, 16, 2, 3}, {128, 128, 8, 1, 3}, {128, 128, 8, 1, 8}};

struct HnW {
    const float* w[7];       // [cin][ks*ks][cout]
    const float* scale[7];   // 1 / sqrt(var + 1e-5)
    const float* shift[7];   // -mean * scale
};

inline size_t hn_walk(const float* base, HnW* out) {
    size_t off = 0;
    HnW w;
    for (int l = 0; l < 7; ++l) {
        const HnLayer& L = kHn[l];
        w.w[l] = base + off; off += (size_t)L.cin * L.ks * L.ks * L.cout;
        w.scale[l] = base + off; off += L.cout;
        w.shift[l] = base + off; off += L.cout;
    }
    off = (off + 63) / 64 * 64;              // the tensor-core blob that follows is 256-byte aligned
    if (out) *out = w;
    return off;
}

// tensor-core path (precision 1), hardnet_tc.cu
size_t hn_tc_blob_floats();
int hn_tc_pack_weights(const HnW& w, float* blob, cudaStream_t st);
size_t hn_tc_workspace_bytes(int n_patches);
int hn_tc_forward(const HnW& w, const float* blob, const float* patches, int n_patches, float* desc, void* workspace,
                  cudaStream_t st, int fp16_operands);
// final 8x8 "valid" layer + BatchNorm + L2 norm on a flat [n][8192] input (hardnet.cu)
int hn_run_final(const float* in, int n, const HnW& w, float* desc, cudaStream_t st);

}  // namespace balf
