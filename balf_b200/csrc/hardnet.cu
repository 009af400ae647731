// HardNet descriptor network for sm_100a (first, CUDA-core fp32 version of row H1).
//
// Reference semantics (restated in oracle/hardnet.py):
//   third_party/hardnet/hardnet_pytorch.py:62-67  input_norm: per-patch (x - mean) / (unbiased std + 1e-7)
//   third_party/hardnet/hardnet_pytorch.py:36-59  7 x [conv (bias=False) -> BatchNorm(affine=False, eval) -> ReLU],
//                                                 no ReLU after the last; Dropout(0.1) is an eval no-op
//   third_party/hardnet/hardnet_pytorch.py:7-15   L2Norm: x / sqrt(sum x^2 + 1e-10)
//   demo/demo_match.py:71-93                      processed in chunks of 1000 patches (per-patch independent in
//                                                 eval mode, so the whole batch is one pass here)
//
// Activations are NCHW fp32 per patch.  Each 3x3 layer is one kernel: a CTA owns one patch, stages its
// zero-haloed input in shared memory and computes register tiles of PX pixels x CH channels with the
// folded BatchNorm + ReLU in the epilogue.  The 8x8 "valid" layer is a [patches x 8192] x [8192 x 128]
// product followed by BatchNorm and the L2 normalisation.
#include "hardnet.cuh"

namespace balf {

static size_t hn_raw_count() {
    size_t n = 0;
    for (int l = 0; l < 7; ++l) n += (size_t)kHn[l].cin * kHn[l].ks * kHn[l].ks * kHn[l].cout + 2 * kHn[l].cout;
    return n;
}

// conv weight [cout][cin][ks][ks] -> [cin][ks*ks][cout]
__global__ void hn_pack_w_kernel(const float* __restrict__ src, int cout, int cin, int taps, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin * taps) return;
    int t = i % taps, ci = (i / taps) % cin, co = i / (taps * cin);
    dst[((size_t)ci * taps + t) * cout + co] = src[i];
}
__global__ void hn_pack_bn_kernel(const float* __restrict__ mean, const float* __restrict__ var, int n,
                                  float* __restrict__ scale, float* __restrict__ shift) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 1.0f / sqrtf(var[i] + 1e-5f);
    scale[i] = s;
    shift[i] = -mean[i] * s;
}

// input_norm: one warp per patch (1024 values); two-pass mean / unbiased std like torch.std
__global__ void hn_input_norm_kernel(const float* __restrict__ x, int n, float* __restrict__ y) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n) return;
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)p * 1024);
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = __ldg(src + lane + 32 * i); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    const float mean = warp_sum(s) * (1.0f / 1024.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float sd = sqrtf(warp_sum(q) * (1.0f / 1023.0f)) + 1e-7f;
    float4* dst = reinterpret_cast<float4*>(y + (size_t)p * 1024);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        dst[lane + 32 * i] = make_float4((v[i].x - mean) / sd, (v[i].y - mean) / sd, (v[i].z - mean) / sd, (v[i].w - mean) / sd);
}

// 3x3 convolution, padding 1, stride S, + folded BN + ReLU.  in [N][CIN][HIN][HIN] -> out [N][COUT][HO][HO].
// Thread tile: PX consecutive output pixels of one row x CH output channels.
template <int CIN, int COUT, int HIN, int S, int PX, int CH>
__global__ void __launch_bounds__(256) hn_conv3_kernel(const float* __restrict__ in, const float* __restrict__ wT,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       float* __restrict__ out) {
    constexpr int HO = HIN / S, HP = HIN + 2, WP = HIN + 2 + ((HIN + 2) % 2 == 0 ? 1 : 0);   // odd row stride
    constexpr int XT = HO / PX;                 // pixel tiles per row
    constexpr int NPT = HO * XT;                // pixel tiles per patch
    constexpr int NCT = COUT / CH;
    extern __shared__ __align__(16) float s_in[];   // [CIN][HP][WP]
    const int n = blockIdx.x;
    for (int i = threadIdx.x; i < CIN * HP * WP; i += blockDim.x) s_in[i] = 0.f;
    __syncthreads();
    const float* src = in + (size_t)n * CIN * HIN * HIN;
    for (int i = threadIdx.x; i < CIN * HIN * HIN; i += blockDim.x) {
        int c = i / (HIN * HIN), r = (i / HIN) % HIN, x = i % HIN;
        s_in[(c * HP + r + 1) * WP + x + 1] = __ldg(src + i);
    }
    __syncthreads();
    for (int task = threadIdx.x; task < NPT * NCT; task += blockDim.x) {
        const int pt = task % NPT, ct = task / NPT;
        const int oy = pt / XT, ox0 = (pt % XT) * PX, co0 = ct * CH;
        float acc[CH][PX];
#pragma unroll
        for (int j = 0; j < CH; ++j)
#pragma unroll
            for (int i = 0; i < PX; ++i) acc[j][i] = 0.f;
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                constexpr int NI = (PX - 1) * S + 3;
                float xin[NI];
                const float* row = s_in + (ci * HP + oy * S + dy) * WP + ox0 * S;
#pragma unroll
                for (int i = 0; i < NI; ++i) xin[i] = row[i];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float* wp = wT + ((size_t)ci * 9 + dy * 3 + dx) * COUT + co0;
                    float w[CH];
#pragma unroll
                    for (int j = 0; j < CH / 4; ++j) *reinterpret_cast<float4*>(&w[4 * j]) = __ldg(reinterpret_cast<const float4*>(wp) + j);
#pragma unroll
                    for (int j = 0; j < CH; ++j)
#pragma unroll
                        for (int i = 0; i < PX; ++i) acc[j][i] = fmaf(w[j], xin[i * S + dx], acc[j][i]);
                }
            }
        }
        float* dst = out + (size_t)n * COUT * HO * HO;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const float sc = __ldg(scale + co0 + j), sh = __ldg(shift + co0 + j);
#pragma unroll
            for (int i = 0; i < PX; ++i) dst[((size_t)(co0 + j) * HO + oy) * HO + ox0 + i] = fmaxf(fmaf(acc[j][i], sc, sh), 0.f);
        }
    }
}

// final layer: out[n][co] = BN(sum_k x[n][k] * wT[k][co]), k = ci*64 + tap (== NCHW flattening of the
// 128 x 8 x 8 input), then L2 normalisation.  A CTA owns PB patches; thread = output channel.
template <int PB>
__global__ void __launch_bounds__(128) hn_final_kernel(const float* __restrict__ in, int n, const float* __restrict__ wT,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       float* __restrict__ desc) {
    constexpr int K = 8192, KC = 512;
    __shared__ float xs[PB][KC];
    __shared__ float red[PB][4];
    const int p0 = blockIdx.x * PB, co = threadIdx.x;
    float acc[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) acc[p] = 0.f;
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        for (int i = threadIdx.x; i < PB * KC; i += 128) {
            int p = i / KC, k = i % KC;
            xs[p][k] = (p0 + p < n) ? __ldg(in + (size_t)(p0 + p) * K + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
            const float w = __ldg(wT + (size_t)(k0 + k) * 128 + co);
#pragma unroll
            for (int p = 0; p < PB; ++p) acc[p] = fmaf(xs[p][k], w, acc[p]);
        }
    }
    const float sc = __ldg(scale + co), sh = __ldg(shift + co);
    float v[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) {
        v[p] = fmaf(acc[p], sc, sh);
        float q = warp_sum(v[p] * v[p]);
        if ((threadIdx.x & 31) == 0) red[p][threadIdx.x >> 5] = q;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PB; ++p) {
        if (p0 + p >= n) break;
        const float nrm = sqrtf(((red[p][0] + red[p][1]) + (red[p][2] + red[p][3])) + 1e-10f);
        desc[(size_t)(p0 + p) * 128 + co] = v[p] / nrm;
    }
}

template <int CIN, int COUT, int HIN, int S, int PX, int CH>
static int hn_launch_conv(const char* name, const float* in, const HnW& w, int layer, float* out, int n, cudaStream_t st) {
    constexpr int HP = HIN + 2, WP = HIN + 2 + ((HIN + 2) % 2 == 0 ? 1 : 0);
    const size_t smem = sizeof(float) * CIN * HP * WP;
    BALF_CUDA_OK(cudaFuncSetAttribute(hn_conv3_kernel<CIN, COUT, HIN, S, PX, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfScope p(name, st);
        hn_conv3_kernel<CIN, COUT, HIN, S, PX, CH><<<n, 256, smem, st>>>(in, w.w[layer], w.scale[layer], w.shift[layer], out);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

int hn_run_final(const float* in, int n, const HnW& w, float* desc, cudaStream_t st) {
    {
        ProfScope p("hn_final", st);
        hn_final_kernel<8><<<cdiv(n, 8), 128, 0, st>>>(in, n, w.w[6], w.scale[6], w.shift[6], desc);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

}  // namespace balf

using namespace balf;

extern "C" int64_t balf_hardnet_raw_weight_count(void) { return (int64_t)hn_raw_count(); }
extern "C" int64_t balf_hardnet_packed_weight_count(void) { return (int64_t)(hn_walk(nullptr, nullptr) + hn_tc_blob_floats()); }

extern "C" int balf_hardnet_pack_weights(const float* raw, float* packed, void* stream) {
    BALF_REQUIRE(raw && packed, "null pointer argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    HnW w;
    hn_walk(packed, &w);
    const float* src = raw;
    for (int l = 0; l < 7; ++l) {
        const HnLayer& L = kHn[l];
        const int taps = L.ks * L.ks, nw = L.cout * L.cin * taps;
        hn_pack_w_kernel<<<cdiv(nw, 256), 256, 0, st>>>(src, L.cout, L.cin, taps, const_cast<float*>(w.w[l]));
        src += nw;
        hn_pack_bn_kernel<<<1, 128, 0, st>>>(src, src + L.cout, L.cout, const_cast<float*>(w.scale[l]), const_cast<float*>(w.shift[l]));
        src += 2 * L.cout;
    }
    BALF_LAUNCH_OK();
    return hn_tc_pack_weights(w, packed + hn_walk(nullptr, nullptr), st);     // second half: tensor-core operand images
}

// fp32 path: two ping-pong activation buffers of N x 32 x 32 x 32 floats; tensor-core path: see hardnet_tc.cu
static size_t hn_fp32_workspace_bytes(int n_patches) { return 2 * align_up((size_t)n_patches * 32 * 1024 * sizeof(float), 256); }
extern "C" size_t balf_hardnet_workspace_bytes(int n_patches, int precision) {
    if (n_patches <= 0) return 0;
    return precision >= 1 ? hn_tc_workspace_bytes(n_patches) : hn_fp32_workspace_bytes(n_patches);
}

extern "C" int balf_hardnet_forward(const float* packed, const float* patches, int n_patches, float* desc, void* workspace,
                                    size_t workspace_bytes, int precision, void* stream) {
    BALF_REQUIRE(packed && patches && desc && workspace, "null pointer argument");
    BALF_REQUIRE(n_patches > 0, "n_patches must be positive");
    BALF_REQUIRE(precision >= 0 && precision <= 2, "precision %d is not built in this library (0 = fp32, 1 = tf32, 2 = fp16 operands)", precision);
    BALF_REQUIRE(workspace_bytes >= balf_hardnet_workspace_bytes(n_patches, precision), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    HnW w;
    const size_t fp32_floats = hn_walk(packed, &w);
    if (precision >= 1) return hn_tc_forward(w, packed + fp32_floats, patches, n_patches, desc, workspace, st, precision == 2);
    float* a = static_cast<float*>(workspace);
    float* b = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up((size_t)n_patches * 32 * 1024 * sizeof(float), 256));
    const int n = n_patches;
    {
        ProfScope p("hn_input_norm", st);
        hn_input_norm_kernel<<<cdiv(n, 8), 256, 0, st>>>(patches, n, a);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    if (int e = hn_launch_conv<1, 32, 32, 1, 8, 8>("hn_conv1", a, w, 0, b, n, st)) return e;
    if (int e = hn_launch_conv<32, 32, 32, 1, 8, 8>("hn_conv2", b, w, 1, a, n, st)) return e;
    if (int e = hn_launch_conv<32, 64, 32, 2, 4, 16>("hn_conv3", a, w, 2, b, n, st)) return e;
    if (int e = hn_launch_conv<64, 64, 16, 1, 4, 16>("hn_conv4", b, w, 3, a, n, st)) return e;
    if (int e = hn_launch_conv<64, 128, 16, 2, 4, 8>("hn_conv5", a, w, 4, b, n, st)) return e;
    if (int e = hn_launch_conv<128, 128, 8, 1, 4, 8>("hn_conv6", b, w, 5, a, n, st)) return e;
    return hn_run_final(a, n, w, desc, st);
}
