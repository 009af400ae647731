// C-ABI plumbing shared by every stage (error string, launch counter) plus the small
// data-movement stages: D0 image -> network input, and the stand-alone depth-to-space.
#include "common.cuh"
#include "../../include/balf_b200.h"

#include <string>

namespace balf {

unsigned long long g_launches = 0;
static thread_local std::string g_error;

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}
const char* last_error() { return g_error.c_str(); }

// D0: demo_match.py:22-29.  float(u8) / 255.f is bit-identical to the reference's float64
// division followed by the fp32 cast for all 256 inputs.
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int C, float* __restrict__ x,
                                     int Hp, int Wp, int top, int left) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y, b = blockIdx.z;
    if (xo >= Wp) return;
    const int yi = yo - top, xi = xo - left;
    const bool in = yi >= 0 && yi < H && xi >= 0 && xi < W;
    const uint8_t* px = img + (((size_t)b * H + (in ? yi : 0)) * W + (in ? xi : 0)) * C;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = in ? (float)px[C == 1 ? 0 : c] / 255.0f : 0.0f;
        x[(((size_t)b * 3 + c) * Hp + yo) * Wp + xo] = v;
    }
}

// tensor_op.py:1-27: out[n, c, r*i+a, r*j+b] = in[n, c*r*r + a*r + b, i, j]
__global__ void pixel_shuffle_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W, int r,
                                     size_t total) {
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int Wo = W * r, Ho = H * r, Co = C / (r * r);
    int xo = (int)(o % Wo);
    size_t t = o / Wo;
    int yo = (int)(t % Ho);
    t /= Ho;
    int co = (int)(t % Co);
    size_t n = t / Co;
    int ci = co * r * r + (yo % r) * r + (xo % r);
    out[o] = in[((n * C + ci) * H + yo / r) * W + xo / r];
}

}  // namespace balf

using namespace balf;

extern "C" const char* balf_last_error(void) { return last_error(); }
extern "C" int balf_abi_version(void) { return BALF_B200_ABI_VERSION; }
extern "C" unsigned long long balf_launch_count(void) { return g_launches; }

extern "C" int balf_pad_geometry(int H, int W, int factor, int* Hp, int* Wp, int* top, int* left) {
    BALF_REQUIRE(H > 0 && W > 0 && factor > 0, "H, W, factor must be positive");
    // make_shape_even (test_utils.py:16-21) then mod_padding_symmetric (test_utils.py:23-32)
    const int he = H + (H & 1), we = W + (W & 1);
    const int ph = he % factor ? ((he + factor) / factor) * factor - he : 0;
    const int pw = we % factor ? ((we + factor) / factor) * factor - we : 0;
    if (Hp) *Hp = he + 2 * (ph / 2);
    if (Wp) *Wp = we + 2 * (pw / 2);
    if (top) *top = ph / 2;
    if (left) *left = pw / 2;
    return 0;
}

extern "C" int balf_preprocess_u8(const uint8_t* img, int B, int H, int W, int C, float* x, int Hp, int Wp, int top,
                                  int left, void* stream) {
    BALF_REQUIRE(img && x, "null pointer argument");
    BALF_REQUIRE(C == 1 || C == 3, "image must have 1 or 3 channels, got %d", C);
    BALF_REQUIRE(B > 0 && H > 0 && W > 0 && top >= 0 && left >= 0 && top + H <= Hp && left + W <= Wp,
                 "image %dx%d at (%d,%d) does not fit the padded size %dx%d", H, W, top, left, Hp, Wp);
    dim3 grid(cdiv(Wp, 128), Hp, B);
    preprocess_u8_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(img, H, W, C, x, Hp, Wp, top, left);
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_pixel_shuffle(const float* in, float* out, int N, int C, int H, int W, int r, void* stream) {
    BALF_REQUIRE(in && out, "null pointer argument");
    BALF_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && r > 0 && C % (r * r) == 0, "bad pixel_shuffle shape");
    size_t total = (size_t)N * C * H * W;
    pixel_shuffle_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, C, H, W,
                                                                                                       r, total);
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}
