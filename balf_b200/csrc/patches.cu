// F1: keypoints -> 32x32 descriptor patches (LAF + scale pyramid + bilinear gather) for sm_100a.
//
// Reference call sites: demo/demo_match.py:62-69
//     lafs = K.feature.laf_from_center_scale_ori(xy, s_mult * ones)          (angle 0)
//     patches = K.feature.extract_patches_from_pyramid(gray / 255., lafs, PS=32)[0]
// kornia 0.7.4 is an un-vendored pip dependency (PARITY UNPINNED, see oracle/thirdparty.py, which
// restates SURVEY.md appendix B2-B3 and is what the tests compare against):
//   * scale = 2 * s_mult / PS is the same for every keypoint, so all of them sample ONE pyramid level
//     L = clamp(floor(log2(scale)), 0, max(0, min(H, W) / PS - 1))  (L = 1 for the demo's 60 / 32);
//   * pyrdown = 5x5 binomial blur ([1 4 6 4 1]^2 / 256, reflect border) followed by a bilinear resize
//     (align_corners = False) to (int(h / 2), int(w // 2));
//   * patch sample (row i, column j) of keypoint (x, y) at a level of size h x w:
//         u_j = (2 j + 1) / PS - 1,   sL = s_mult * min(h - 1, w - 1) / min(H - 1, W - 1)
//         gx = sL * u_j + x * (w - 1) / (W - 1)         (pixel units of the level)
//         ix = gx * w / (w - 1) - 0.5                   (grid normalisation + grid_sample unnormalise)
//     clamped to [0, w - 1] (padding_mode = 'border'), bilinear.
//
// This is gather work bound by L2 / HBM latency: the level image (<= 1 MB) stays L2 resident, one CTA
// writes one patch (4 KB) with coalesced 128-byte rows.
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

__device__ __forceinline__ int reflect_idx(int i, int n) {          // 'reflect' (no edge repeat), n >= 2
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);
}

template <typename T> __device__ __forceinline__ float px_load(const T* p);
template <> __device__ __forceinline__ float px_load<uint8_t>(const uint8_t* p) { return (float)__ldg(p) / 255.0f; }
template <> __device__ __forceinline__ float px_load<float>(const float* p) { return __ldg(p); }

// 5x5 binomial blur at (y, x) of an h x w image with reflect border
template <typename T>
__device__ __forceinline__ float blur5(const T* img, int h, int w, int y, int x) {
    const float k[5] = {1.f, 4.f, 6.f, 4.f, 1.f};
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const T* row = img + (size_t)reflect_idx(y + i - 2, h) * w;
#pragma unroll
        for (int j = 0; j < 5; ++j) acc = fmaf(px_load<T>(row + reflect_idx(x + j - 2, w)), k[i] * k[j] * (1.0f / 256.0f), acc);
    }
    return acc;
}

// one pyramid step: in [B, h, w] -> out [B, ho, wo], ho = int(h / 2), wo = w / 2
template <typename T>
__global__ void pyrdown_kernel(const T* __restrict__ in, int h, int w, float* __restrict__ out, int ho, int wo) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= wo) return;
    const T* img = in + (size_t)b * h * w;
    // F.interpolate(mode='bilinear', align_corners=False): src = max((dst + 0.5) * (in / out) - 0.5, 0)
    const float sy = fmaxf(((float)y + 0.5f) * ((float)h / (float)ho) - 0.5f, 0.f);
    const float sx = fmaxf(((float)x + 0.5f) * ((float)w / (float)wo) - 0.5f, 0.f);
    const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float v00 = blur5(img, h, w, y0, x0), v01 = blur5(img, h, w, y0, x1);
    const float v10 = blur5(img, h, w, y1, x0), v11 = blur5(img, h, w, y1, x1);
    out[((size_t)b * ho + y) * wo + x] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// one CTA per keypoint, PS*PS samples; kp [B, K, 2] fp32 (x, y) in full-resolution pixel units
template <typename T>
__global__ void __launch_bounds__(256) patch_gather_kernel(const T* __restrict__ level, int h, int w, int H, int W,
                                                           const float* __restrict__ kp, const int32_t* __restrict__ count,
                                                           int K, float s_mult, int PS, float* __restrict__ patches) {
    const int k = blockIdx.x, b = blockIdx.y;
    if (count && k >= count[b]) return;
    const T* img = level + (size_t)b * h * w;
    const float x = __ldg(kp + ((size_t)b * K + k) * 2), y = __ldg(kp + ((size_t)b * K + k) * 2 + 1);
    const float m0 = (float)min(H - 1, W - 1), ml = (float)min(h - 1, w - 1);
    const float sL = __fmul_rn(__fdiv_rn(s_mult, m0), ml);
    const float tx = __fmul_rn(__fdiv_rn(x, (float)(W - 1)), (float)(w - 1)), ty = __fmul_rn(__fdiv_rn(y, (float)(H - 1)), (float)(h - 1));
    float* dst = patches + ((size_t)b * K + k) * PS * PS;
    for (int i = threadIdx.x; i < PS * PS; i += blockDim.x) {
        const int r = i / PS, c = i - r * PS;
        // every step rounded separately (no FMA contraction), in the order the torch CPU path evaluates
        // affine_grid -> grid normalisation -> grid_sample's unnormalise: (g + 1) * (size / 2) - 0.5
        const float uc = __fsub_rn(__fdiv_rn(__fadd_rn(2.0f * (float)c, 1.0f), (float)PS), 1.0f);
        const float ur = __fsub_rn(__fdiv_rn(__fadd_rn(2.0f * (float)r, 1.0f), (float)PS), 1.0f);
        const float px = __fadd_rn(__fmul_rn(sL, uc), tx), py = __fadd_rn(__fmul_rn(sL, ur), ty);
        const float gx = __fsub_rn(__fdiv_rn(2.0f * px, (float)(w - 1)), 1.0f);
        const float gy = __fsub_rn(__fdiv_rn(2.0f * py, (float)(h - 1)), 1.0f);
        float ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)w * 0.5f), 0.5f);
        float iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)h * 0.5f), 0.5f);
        ix = fminf(fmaxf(ix, 0.f), (float)(w - 1));
        iy = fminf(fmaxf(iy, 0.f), (float)(h - 1));
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float ax = ix - fx, ay = iy - fy;
        const float v00 = px_load<T>(img + (size_t)y0 * w + x0);
        const float v01 = x1 < w ? px_load<T>(img + (size_t)y0 * w + x1) : 0.f;
        const float v10 = y1 < h ? px_load<T>(img + (size_t)y1 * w + x0) : 0.f;
        const float v11 = (x1 < w && y1 < h) ? px_load<T>(img + (size_t)y1 * w + x1) : 0.f;
        dst[i] = v00 * (1.f - ax) * (1.f - ay) + v01 * ax * (1.f - ay) + v10 * (1.f - ax) * ay + v11 * ax * ay;
    }
}

static int pyramid_level(int H, int W, float s_mult, int PS) {
    // kornia: scale = 2 * sqrt(|det A| + 1e-10) / PS on the denormalised LAF, A = s_mult * I
    const float scale = 2.0f * sqrtf(fabsf(s_mult * s_mult) + 1e-10f) / (float)PS;
    int lvl = (int)floorf(fmaxf(log2f(scale), 0.f));
    const int max_level = (H < W ? H : W) / PS;
    const int hi = max_level - 1 > 0 ? max_level - 1 : 0;
    return lvl < hi ? lvl : hi;
}

}  // namespace balf

using namespace balf;

extern "C" int balf_patch_pyramid_level(int H, int W, float s_mult, int PS) { return pyramid_level(H, W, s_mult, PS); }

// two ping-pong level buffers (level 1 is the largest: (H/2) x (W/2) floats per image)
extern "C" size_t balf_patches_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return 2 * align_up(sizeof(float) * (size_t)B * (H / 2 + 1) * (W / 2 + 1), 256);
}

extern "C" int balf_extract_patches_u8(const uint8_t* gray, int B, int H, int W, const float* kpts, const int32_t* count,
                                       int K, float s_mult, int PS, float* patches, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    BALF_REQUIRE(gray && kpts && patches && workspace, "null pointer argument");
    BALF_REQUIRE(B > 0 && H >= 2 && W >= 2 && K > 0 && PS > 0 && PS <= 64, "bad patch extraction shape (B=%d H=%d W=%d K=%d PS=%d)", B, H, W, K, PS);
    BALF_REQUIRE(workspace_bytes >= balf_patches_workspace_bytes(B, H, W), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int level = pyramid_level(H, W, s_mult, PS);
    float* buf[2];
    buf[0] = static_cast<float*>(workspace);
    buf[1] = reinterpret_cast<float*>(static_cast<char*>(workspace) + balf_patches_workspace_bytes(B, H, W) / 2);
    int h = H, w = W;
    const float* cur = nullptr;
    for (int l = 0; l < level; ++l) {
        const int ho = (int)((float)h / 2.0f), wo = w / 2;
        BALF_REQUIRE(ho >= 2 && wo >= 2, "image too small for pyramid level %d", level);
        dim3 grid(cdiv(wo, 128), ho, B);
        ProfScope p("patch_pyrdown", st);
        if (l == 0) pyrdown_kernel<uint8_t><<<grid, 128, 0, st>>>(gray, h, w, buf[0], ho, wo);
        else pyrdown_kernel<float><<<grid, 128, 0, st>>>(cur, h, w, buf[l & 1], ho, wo);
        BALF_COUNT_LAUNCH(1);
        cur = buf[l & 1];
        h = ho;
        w = wo;
    }
    BALF_LAUNCH_OK();
    dim3 grid(K, B);
    {
        ProfScope p("patch_gather", st);
        if (level == 0) patch_gather_kernel<uint8_t><<<grid, 256, 0, st>>>(gray, h, w, H, W, kpts, count, K, s_mult, PS, patches);
        else patch_gather_kernel<float><<<grid, 256, 0, st>>>(cur, h, w, H, W, kpts, count, K, s_mult, PS, patches);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}
