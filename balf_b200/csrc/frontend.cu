// SURVEY.md section 8(f) rows behind the C-ABI: the image front end (PIL "L" conversion, float-image padding), the
// multi-scale pyramid front end (bilinear down-scaling fused with the /255, pad and HWC->CHW of row D0) and the
// score-ordered merge of per-level keypoint lists.  All HBM-bound elementwise / gather kernels: one thread per output
// element, consecutive threads on consecutive addresses.
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

// demo/demo_match.py:13-19: PIL Image.convert('L') = ITU-R 601-2 luma in 16.16 fixed point
__global__ void rgb_to_gray_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    gray[i] = (uint8_t)((19595u * r + 38470u * g + 7471u * b + 32768u) >> 16);
}

// balf/utils/train_utils.py:420-430: an already normalised float image [B,H,W,C] -> padded CHW network input
__global__ void preprocess_f32_kernel(const float* __restrict__ img, int H, int W, int C, float* __restrict__ x,
                                      int Hp, int Wp, int top, int left) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y, b = blockIdx.z;
    if (xo >= Wp) return;
    const int yi = yo - top, xi = xo - left;
    const bool in = yi >= 0 && yi < H && xi >= 0 && xi < W;
    const float* px = img + (((size_t)b * H + (in ? yi : 0)) * W + (in ? xi : 0)) * C;
#pragma unroll
    for (int c = 0; c < 3; ++c) x[(((size_t)b * 3 + c) * Hp + yo) * Wp + xo] = in ? px[C == 1 ? 0 : c] : 0.0f;
}

// Pyramid level: bilinear resize of the uint8 image to Hs x Ws (half-pixel centres, no antialiasing: the arithmetic of
// torch.nn.functional.interpolate(mode='bilinear', align_corners=False)), then /255, zero pad, HWC -> CHW.
// Every product / sum is a separately rounded fp32 operation (no FMA contraction) so that the CPU restatement (oracle/multiscale.py) reproduces
// it bit for bit.
__device__ __forceinline__ void src_index(int d, float scale, int n, int& i0, int& i1, float& l1) {
    float s = __fsub_rn(__fmul_rn(__fadd_rn((float)d, 0.5f), scale), 0.5f);
    s = s < 0.0f ? 0.0f : s;
    i0 = (int)s;
    if (i0 > n - 1) i0 = n - 1;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
    l1 = __fsub_rn(s, (float)i0);
}
__global__ void resize_preprocess_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int C, int Hs, int Ws, float sy,
                                            float sx, float* __restrict__ x, int Hp, int Wp, int top, int left) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y, b = blockIdx.z;
    if (xo >= Wp) return;
    const int yd = yo - top, xd = xo - left;
    const bool in = yd >= 0 && yd < Hs && xd >= 0 && xd < Ws;
    float out[3] = {0.f, 0.f, 0.f};
    if (in) {
        int y0, y1, x0, x1;
        float ly, lx;
        src_index(yd, sy, H, y0, y1, ly);
        src_index(xd, sx, W, x0, x1, lx);
        const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
        const uint8_t* base = img + (size_t)b * H * W * C;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int cc = C == 1 ? 0 : c;
            const float p00 = base[((size_t)y0 * W + x0) * C + cc], p01 = base[((size_t)y0 * W + x1) * C + cc];
            const float p10 = base[((size_t)y1 * W + x0) * C + cc], p11 = base[((size_t)y1 * W + x1) * C + cc];
            const float a = __fadd_rn(__fmul_rn(hx, p00), __fmul_rn(lx, p01));
            const float d = __fadd_rn(__fmul_rn(hx, p10), __fmul_rn(lx, p11));
            out[c] = __fdiv_rn(__fadd_rn(__fmul_rn(hy, a), __fmul_rn(ly, d)), 255.0f);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) x[(((size_t)b * 3 + c) * Hp + yo) * Wp + xo] = out[c];
}

// Merge of L per-level lists, each already ordered (score descending, raster ascending): the global rank of entry i of
// level l is i + sum over the other levels m of the number of their entries that precede it -- score greater, or equal
// with m < l -- found by binary search.  One thread per entry, no sorting, deterministic.  Coordinates return to the
// level-0 frame by the inverse of the half-pixel mapping of the resize.
struct MergeLevels {
    const int32_t* xy[BALF_MAX_LEVELS];
    const float* score[BALF_MAX_LEVELS];
    const int32_t* count[BALF_MAX_LEVELS];
    float sx[BALF_MAX_LEVELS], sy[BALF_MAX_LEVELS];      // level-0 pixels per level pixel
    int n;
};
__global__ void merge_levels_kernel(MergeLevels lv, int K, int Kout, float* __restrict__ xy_out, float* __restrict__ score_out,
                                    int32_t* __restrict__ level_out, int32_t* __restrict__ count_out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= lv.n * K) return;
    const int l = e / K, i = e - l * K;
    if (e == 0) {
        int tot = 0;
        for (int m = 0; m < lv.n; ++m) tot += lv.count[m][b];
        count_out[b] = tot < Kout ? tot : Kout;
    }
    if (i >= lv.count[l][b]) return;
    const float s = lv.score[l][(size_t)b * K + i];
    int rank = i;
    for (int m = 0; m < lv.n; ++m) {
        if (m == l) continue;
        const float* sm = lv.score[m] + (size_t)b * K;
        int lo = 0, hi = lv.count[m][b];                 // first index whose entry does NOT precede (s, l)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const float v = sm[mid];
            const bool before = v > s || (v == s && m < l);
            if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    if (rank >= Kout) return;
    const int32_t* p = lv.xy[l] + ((size_t)b * K + i) * 2;
    const size_t o = (size_t)b * Kout + rank;
    xy_out[2 * o] = __fsub_rn(__fmul_rn(__fadd_rn((float)p[0], 0.5f), lv.sx[l]), 0.5f);
    xy_out[2 * o + 1] = __fsub_rn(__fmul_rn(__fadd_rn((float)p[1], 0.5f), lv.sy[l]), 0.5f);
    score_out[o] = s;
    level_out[o] = l;
}

}  // namespace balf

using namespace balf;

extern "C" int balf_rgb_to_gray_u8(const uint8_t* rgb, int B, int H, int W, uint8_t* gray, void* stream) {
    BALF_REQUIRE(rgb && gray, "null pointer argument");
    BALF_REQUIRE(B > 0 && H > 0 && W > 0, "B, H, W must be positive");
    const size_t n = (size_t)B * H * W;
    rgb_to_gray_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb, gray, n);
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_preprocess_f32(const float* img, int B, int H, int W, int C, float* x, int Hp, int Wp, int top, int left,
                                   void* stream) {
    BALF_REQUIRE(img && x, "null pointer argument");
    BALF_REQUIRE(C == 1 || C == 3, "image must have 1 or 3 channels, got %d", C);
    BALF_REQUIRE(B > 0 && H > 0 && W > 0 && top >= 0 && left >= 0 && top + H <= Hp && left + W <= Wp,
                 "image %dx%d at (%d,%d) does not fit the padded size %dx%d", H, W, top, left, Hp, Wp);
    dim3 grid(cdiv(Wp, 128), Hp, B);
    preprocess_f32_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(img, H, W, C, x, Hp, Wp, top, left);
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_resize_preprocess_u8(const uint8_t* img, int B, int H, int W, int C, int Hs, int Ws, float* x, int Hp,
                                         int Wp, int top, int left, void* stream) {
    BALF_REQUIRE(img && x, "null pointer argument");
    BALF_REQUIRE(C == 1 || C == 3, "image must have 1 or 3 channels, got %d", C);
    BALF_REQUIRE(B > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0 && top >= 0 && left >= 0 && top + Hs <= Hp && left + Ws <= Wp,
                 "level %dx%d at (%d,%d) does not fit the padded size %dx%d", Hs, Ws, top, left, Hp, Wp);
    dim3 grid(cdiv(Wp, 128), Hp, B);
    {
        ProfScope p("resize_preprocess_u8", static_cast<cudaStream_t>(stream));
        resize_preprocess_u8_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(img, H, W, C, Hs, Ws, (float)H / (float)Hs,
                                                                                       (float)W / (float)Ws, x, Hp, Wp, top, left);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_merge_levels_topk(int n_levels, const int32_t* const* xy, const float* const* score, const int32_t* const* count,
                                      const float* scale_x, const float* scale_y, int B, int K, int k_out, float* xy_out,
                                      float* score_out, int32_t* level_out, int32_t* count_out, void* stream) {
    BALF_REQUIRE(n_levels >= 1 && n_levels <= BALF_MAX_LEVELS, "1 .. %d pyramid levels (got %d)", BALF_MAX_LEVELS, n_levels);
    BALF_REQUIRE(xy && score && count && scale_x && scale_y && xy_out && score_out && level_out && count_out, "null pointer argument");
    BALF_REQUIRE(B > 0 && K > 0 && k_out > 0, "B, K, k_out must be positive");
    MergeLevels lv;
    lv.n = n_levels;
    for (int l = 0; l < n_levels; ++l) {
        BALF_REQUIRE(xy[l] && score[l] && count[l], "null list pointer at level %d", l);
        lv.xy[l] = xy[l]; lv.score[l] = score[l]; lv.count[l] = count[l];
        lv.sx[l] = scale_x[l]; lv.sy[l] = scale_y[l];
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BALF_CUDA_OK(cudaMemsetAsync(xy_out, 0, (size_t)B * k_out * 2 * sizeof(float), st));
    BALF_CUDA_OK(cudaMemsetAsync(score_out, 0, (size_t)B * k_out * sizeof(float), st));
    BALF_CUDA_OK(cudaMemsetAsync(level_out, 0, (size_t)B * k_out * sizeof(int32_t), st));
    dim3 grid(cdiv(n_levels * K, 256), B);
    {
        ProfScope p("merge_levels", st);
        merge_levels_kernel<<<grid, 256, 0, st>>>(lv, K, k_out, xy_out, score_out, level_out, count_out);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}
