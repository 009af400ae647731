// HardNet on the sm_100a tensor cores: precision 1 ("tf32") of balf_hardnet_forward.
//
// Reference semantics: third_party/hardnet/hardnet_pytorch.py:36-59 (7 x conv + BatchNorm(affine=False, eval)
// (+ ReLU)), :62-67 input_norm, :7-15 L2Norm -- restated in oracle/hardnet.py.
//
// The five 3x3 convolutions after the first (99 % of the 78.2 MFLOP per patch) run as implicit GEMMs on tcgen05
// (kind::tf32, fp32 accumulation in TMEM), the final 8x8 "valid" layer as a split-K [patches x 8192] x [8192 x 128]
// product on the same pipe; only the first layer (K = 9) stays on the CUDA cores.
//
// Implicit GEMM without im2col.  A layer's input lives in shared memory as the canonical K-major SWIZZLE_NONE
// operand ("chunk-major": [C/4 chunks][R rows][4 floats], one row per position of a zero-haloed pixel grid), so the
// A operand of tap (dy, dx) is the SAME buffer viewed with its start address shifted by (dy-1) * pitch + (dx-1)
// rows: nine shifted views x C/8 K-steps accumulate into one TMEM tile of 128 consecutive grid positions.  Outputs at
// halo positions are computed and thrown away (6-20 % of the rows).  Stride-2 layers read four parity planes
// (even/odd rows x even/odd columns, written that way by the previous layer's epilogue), which turns the stride into
// a plane choice plus a 0/-1 row/column shift.
//
// Activations travel between layers through HBM in exactly the shared-memory image of the consumer (tf32-rounded),
// so a layer's input is ONE TMA bulk copy per patch; the consumer zero-fills the halo rows in shared memory.
// BatchNorm is folded (scale into the weights when packing, shift added in the epilogue with the ReLU).
// Weights stream through a 3-slot ring of pre-packed [Cout x KB] blocks; G patches share every block.
#include <cuda_fp16.h>
#include "hardnet.cuh"
#include "umma.cuh"

namespace balf {
using namespace umma;

constexpr int kHnSlots = 3;
constexpr int kHnThreads = 128;
// conv kernels: two warp groups of 128 threads share the epilogue (tile idx = g * T + t goes to group idx % 2): the layers spend
// most of their time in the TMEM -> ReLU -> fp16 -> next-layer-image epilogue (tensor pipe 16-39 % busy with one group)
constexpr int kHnConvThreads = 256;
// final layer (8x8 "valid" conv = [patches x 8192] x [8192 x 128]): K blocks of 32 = (position, 32 channels)
constexpr int kFinalKB = 32, kFinalBlocks = 8192 / kFinalKB, kFinalBlockFloats = 128 * kFinalKB, kFinalSplit = 4;

__host__ __device__ constexpr int hn_cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int hn_cmin(int a, int b) { return a < b ? a : b; }

// geometry of one tensor-core layer: HOUT x HOUT outputs from a (STRIDE * HOUT)^2 input
// EB = bytes per operand element: 4 = tf32 (4 per 16-byte chunk, K = 8 per MMA), 2 = fp16 (8 per chunk, K = 16 per MMA)
template <int CIN, int COUT, int STRIDE, int HOUT, int G, int KB, int EB = 4>
struct HnGeom {
    static constexpr int EPC = 16 / EB;                                // elements per 16-byte chunk
    static constexpr int KSTEP = 32 / EB;                              // K per MMA (two chunks)
    static constexpr int PW = STRIDE == 1 ? HOUT + 2 : HOUT + 1;      // row pitch of the position grid
    static constexpr int PH = PW;
    static constexpr int PSZ = PW * PH;                                // positions per plane
    static constexpr int NPL = STRIDE == 1 ? 1 : 4;                    // parity planes
    static constexpr int Q0 = PW + 1, QLAST = HOUT * PW + HOUT;        // first / last output position
    static constexpr int NPOS = QLAST - Q0 + 1;
    static constexpr int T = (NPOS + 127) / 128;                       // 128-row accumulator tiles per patch
    static constexpr int MAXROW = (NPL - 1) * PSZ + (T == 1 ? Q0 + 127 : QLAST) + (STRIDE == 1 ? PW + 1 : 0);
    static constexpr int R = (hn_cmax(NPL * PSZ, MAXROW + 1) + 7) / 8 * 8;   // rows per K-chunk plane
    static constexpr int NKP = CIN / KB, NBLK = 9 * NKP;               // weight blocks: (tap, K part)
    static constexpr uint32_t patch_bytes = (uint32_t)CIN * R * EB;
    static constexpr uint32_t block_bytes = (uint32_t)COUT * KB * EB;
    static constexpr int ncols_need = G * T * COUT;
    static constexpr int ncols = ncols_need <= 32 ? 32 : ncols_need <= 64 ? 64 : ncols_need <= 128 ? 128 : ncols_need <= 256 ? 256 : 512;
    static constexpr size_t smem = (size_t)G * patch_bytes + (size_t)kHnSlots * block_bytes + COUT * 4 + 128;
    static_assert(ncols_need <= 512, "accumulators exceed tensor memory");
    // first position of tile t (the last tile is pulled back so that it ends at QLAST: no rows past the grid are read)
    __host__ __device__ static constexpr int qs(int t) { return T == 1 ? Q0 : hn_cmin(Q0 + 128 * t, QLAST + 1 - 128); }
    // row offset of tap (dy, dx) relative to the output position
    __host__ __device__ static constexpr int tap_shift(int dy, int dx) {
        if (STRIDE == 1) return (dy - 1) * PW + (dx - 1);
        const int py = dy == 1 ? 0 : 1, px = dx == 1 ? 0 : 1;
        return (py * 2 + px) * PSZ + (dy == 0 ? -PW : 0) + (dx == 0 ? -1 : 0);
    }
};

// high word of a SWIZZLE_NONE descriptor (SBO = 128, version 1); low word = (LBO >> 4) << 16 | addr >> 4
constexpr uint32_t kHnDescHi = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint64_t hn_desc(uint32_t addr, uint32_t lbo) {
    return ((uint64_t)kHnDescHi << 32) | (uint64_t)(((lbo >> 4) << 16) | ((addr >> 4) & 0x3FFFu));
}

// destination row of output pixel (oy, ox) in the next layer's input image.  NEXT: 0 = one haloed grid (next layer has
// stride 1), 1 = four parity planes (next layer has stride 2)
template <int NEXT, int HOUT>
__device__ __forceinline__ int hn_next_row(int oy, int ox) {
    if (NEXT == 0) return (oy + 1) * (HOUT + 2) + ox + 1;
    constexpr int PWN = HOUT / 2 + 1;
    return ((oy & 1) * 2 + (ox & 1)) * (PWN * PWN) + ((oy >> 1) + 1) * PWN + (ox >> 1) + 1;
}

// 8 consecutive channels of one position -> 16 bytes of fp16 (round to nearest, saturating)
__device__ __forceinline__ uint4 hn_pack8(const float* v) {
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[e]) : "f"(v[2 * e + 1]), "f"(v[2 * e]));
    return make_uint4(h[0], h[1], h[2], h[3]);
}
__host__ __device__ constexpr uint32_t hn_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void hn_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

template <int CIN, int COUT, int STRIDE, int HOUT, int G, int KB, int NEXT, int RN, int EB>
__global__ void __launch_bounds__(kHnConvThreads, 1)
hn_tc_conv_kernel(const float* __restrict__ in_, const float* __restrict__ wblk_, const float* __restrict__ shift,
                  float* __restrict__ out_, int n) {
    using Ge = HnGeom<CIN, COUT, STRIDE, HOUT, G, KB, EB>;
    constexpr int EPC = Ge::EPC;
    const unsigned char* in = reinterpret_cast<const unsigned char*>(in_);
    const unsigned char* wblk = reinterpret_cast<const unsigned char*>(wblk_);
    unsigned char* out = reinterpret_cast<unsigned char*>(out_);
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_in = reinterpret_cast<float*>(smem);
    const uint32_t in_addr = smem_u32(smem), ring_addr = in_addr + G * Ge::patch_bytes;
    float* s_shift = reinterpret_cast<float*>(smem + (size_t)G * Ge::patch_bytes + (size_t)kHnSlots * Ge::block_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + COUT);
    uint64_t* full = bars;                       // [kHnSlots]
    uint64_t* empty = bars + kHnSlots;           // [kHnSlots]
    uint64_t* in_full = bars + 2 * kHnSlots;
    uint64_t* done = bars + 2 * kHnSlots + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kHnSlots + 2);
    const int tid = threadIdx.x, row = tid & 127, wg = tid >> 7;     // accumulator row / epilogue warp group of this thread
    const bool w0 = warp0_uniform();

    if (tid < 32) tmem_alloc(tmem_slot, Ge::ncols);
    if (tid == 0) {
        for (int i = 0; i < kHnSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(in_full, 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < COUT; i += kHnConvThreads) s_shift[i] = __ldg(shift + i);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *tmem_slot;
    const uint32_t lane_base = tm + ((uint32_t)(row & ~31) << 16);

    const int ngroups = (n + G - 1) / G;
    const int my_groups = (int)blockIdx.x < ngroups ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // ring state (only meaningful in the elected lane of warp 0)
    uint32_t pcnt = 0, ccnt = 0, pb = 0, to_load = (uint32_t)Ge::NBLK * (uint32_t)my_groups;
    auto ring_top_up = [&]() {
        while (to_load > 0 && pcnt < ccnt + kHnSlots) {
            const uint32_t slot = pcnt % kHnSlots, use = pcnt / kHnSlots;
            if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
            mbar_expect_tx(&full[slot], Ge::block_bytes);
            bulk_g2s(smem + (size_t)G * Ge::patch_bytes + (size_t)slot * Ge::block_bytes, wblk + (size_t)pb * Ge::block_bytes, Ge::block_bytes, &full[slot]);
            ++pcnt; --to_load;
            if (++pb == (uint32_t)Ge::NBLK) pb = 0;
        }
    };
    auto load_group = [&](int grp) {
        const int p0 = grp * G, valid = hn_cmin(G, n - p0);
        mbar_expect_tx(in_full, (uint32_t)valid * Ge::patch_bytes);
        for (int g = 0; g < valid; ++g)
            bulk_g2s(smem + (size_t)g * Ge::patch_bytes, in + (size_t)(p0 + g) * Ge::patch_bytes, Ge::patch_bytes, in_full);
    };
    if (w0 && elect_one() && my_groups > 0) { load_group(blockIdx.x); ring_top_up(); }

    uint32_t in_phase = 0, done_phase = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        // ---- input image(s) landed; zero-fill the halo rows
        if (tid == 0) mbar_wait(in_full, in_phase & 1);
        ++in_phase;
        __syncthreads();
        {
            constexpr int NH = Ge::NPL == 1 ? 2 * Ge::PW + 2 * (Ge::PH - 2) : 4 * (Ge::PW + Ge::PH - 1);
            for (int i = tid; i < NH * (CIN / EPC) * G; i += kHnConvThreads) {
                const int h = i % NH, c = (i / NH) % (CIN / EPC), g = i / (NH * (CIN / EPC));
                int row;
                if (Ge::NPL == 1) {
                    if (h < Ge::PW) row = h;
                    else if (h < 2 * Ge::PW) row = (Ge::PH - 1) * Ge::PW + (h - Ge::PW);
                    else { const int k = h - 2 * Ge::PW; row = (1 + (k >> 1)) * Ge::PW + ((k & 1) ? Ge::PW - 1 : 0); }
                } else {
                    const int pl = h / (Ge::PW + Ge::PH - 1), k = h - pl * (Ge::PW + Ge::PH - 1);
                    row = pl * Ge::PSZ + (k < Ge::PW ? k : (k - Ge::PW + 1) * Ge::PW);
                }
                *reinterpret_cast<float4*>(s_in + (size_t)g * (Ge::patch_bytes / 4) + ((size_t)c * Ge::R + row) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        // ---- MMAs: nine shifted views x K parts, every weight block shared by the G patches and their tiles
        if (w0 && elect_one()) {
            constexpr uint32_t idesc = EB == 2 ? hn_idesc_f16(128, COUT) : make_idesc_tf32(128, COUT);
            constexpr uint32_t a_lbo = (uint32_t)Ge::R * 16u, b_lbo = (uint32_t)COUT * 16u;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap - dy * 3;
                const int sh = Ge::tap_shift(dy, dx);
#pragma unroll
                for (int kp = 0; kp < Ge::NKP; ++kp) {
                    ring_top_up();
                    const uint32_t slot = ccnt % kHnSlots;
                    mbar_wait(&full[slot], (ccnt / kHnSlots) & 1);
                    fence_after_sync();
                    const uint64_t b_desc = hn_desc(ring_addr + slot * Ge::block_bytes, b_lbo);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
#pragma unroll
                        for (int t = 0; t < Ge::T; ++t) {
                            const uint64_t a_desc = hn_desc(in_addr + (uint32_t)g * Ge::patch_bytes + (uint32_t)(Ge::qs(t) + sh) * 16u, a_lbo);
#pragma unroll
                            for (int k8 = 0; k8 < KB / Ge::KSTEP; ++k8) {
                                const uint32_t kc = (uint32_t)(kp * KB + k8 * Ge::KSTEP) / (uint32_t)EPC;
                                if (EB == 2)
                                    hn_mma_f16(tm + (uint32_t)((g * Ge::T + t) * COUT), a_desc + ((kc * a_lbo) >> 4),
                                               b_desc + (((uint32_t)k8 * 2u * b_lbo) >> 4), idesc, !(tap == 0 && kp == 0 && k8 == 0));
                                else
                                    mma_tf32(tm + (uint32_t)((g * Ge::T + t) * COUT), a_desc + ((kc * a_lbo) >> 4),
                                             b_desc + (((uint32_t)k8 * 2u * b_lbo) >> 4), idesc, !(tap == 0 && kp == 0 && k8 == 0));
                            }
                        }
                    }
                    commit(&empty[slot]);
                    ++ccnt;
                }
            }
            commit(done);
        }
        if (tid == 0) mbar_wait(done, done_phase & 1);
        ++done_phase;
        __syncthreads();
        fence_after_sync();
        // ---- the input buffers are free: start the next group's copies under the epilogue
        if (w0 && elect_one() && grp + (int)gridDim.x < ngroups) load_group(grp + gridDim.x);
        // ---- epilogue: + shift, ReLU, round, scatter into the next layer's image
        const int p0 = grp * G;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (p0 + g >= n) break;                         // uniform over the CTA
#pragma unroll
            for (int t = 0; t < Ge::T; ++t) {
                if (((g * Ge::T + t) & 1) != wg) continue;          // this tile belongs to the other warp group (uniform per warp)
                const int q = Ge::qs(t) + row;
                const int gy = q / Ge::PW, gx = q - gy * Ge::PW;
                const int oy = gy - 1, ox = gx - 1;
                bool valid = q <= Ge::QLAST && ox >= 0 && ox < HOUT && oy >= 0 && oy < HOUT;
                if (t > 0 && q < Ge::qs(t - 1) + 128) valid = false;     // the pulled-back last tile repeats rows
                // byte address of this position's first chunk in the consumer's image; consecutive chunks are `cstride` bytes apart
                unsigned char* dst;
                size_t cstride;
                if (NEXT == 2) {
                    // final layer's A operand: [tile of 128 patches][K block = (position, 32 channels)][32 / EPC chunks][128 rows][16 B]
                    const int p = p0 + g;
                    dst = out + (((size_t)(p >> 7) * kFinalBlocks + (size_t)(oy * 8 + ox) * (COUT / 32)) * kFinalBlockFloats) * EB + (size_t)(p & 127) * 16;
                    cstride = 128 * 16;
                } else {
                    dst = out + (size_t)(p0 + g) * ((size_t)COUT * RN * EB) + (size_t)hn_next_row<NEXT == 1 ? 1 : 0, HOUT>(oy, ox) * 16;
                    cstride = (size_t)RN * 16;
                }
#pragma unroll
                for (int c0 = 0; c0 < COUT; c0 += 32) {
                    float v[32];
                    tmem_ld32(lane_base + (uint32_t)((g * Ge::T + t) * COUT + c0), v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + s_shift[c0 + i], 0.f);
                        // the 32 channels of this load are one K block of the final layer (NEXT == 2: chunk index restarts per
                        // block, blocks are kFinalBlockFloats elements apart) or 32 / EPC consecutive chunks of the next image
                        unsigned char* d0 = NEXT == 2 ? dst + (size_t)(c0 / 32) * kFinalBlockFloats * EB : dst + (size_t)(c0 / EPC) * cstride;
                        if (EB == 2) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(d0 + (size_t)j * cstride) = hn_pack8(v + 8 * j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4*>(d0 + (size_t)j * cstride) = to_tf32(make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                        }
                    }
                }
            }
        }
        fence_before_sync();
        __syncthreads();                                    // TMEM reads retired before the next group's MMAs
        fence_after_sync();
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, Ge::ncols);
}

// ------------------------------------------------------------------------------------------ first layer (CUDA cores)
// input_norm (per-patch mean, unbiased std + 1e-7) + conv 1 -> 32 (3x3, pad 1) + folded BatchNorm + ReLU, written as
// the haloed chunk-major image conv2 reads.  One CTA per patch.
constexpr int kL1R = HnGeom<32, 32, 1, 32, 1, 32>::R;        // 1160 (34 x 34 grid positions, rounded up to 8 rows)
template <int EB>
__global__ void __launch_bounds__(256) hn_tc_first_kernel(const float* __restrict__ x, const float* __restrict__ w9,
                                                          const float* __restrict__ shift, float* __restrict__ out) {
    __shared__ float xs[34][35];
    __shared__ float wsm[9][32];
    __shared__ float ssh[32];
    __shared__ float red[8];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 34 * 35; i += 256) (&xs[0][0])[i] = 0.f;
    for (int i = tid; i < 288; i += 256) (&wsm[0][0])[i] = __ldg(w9 + i);
    if (tid < 32) ssh[tid] = __ldg(shift + tid);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)p * 1024) + tid);
    float s = warp_sum((v.x + v.y) + (v.z + v.w));
    if (lane == 0) red[warp] = s;
    __syncthreads();
    const float mean = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) * (1.0f / 1024.0f);
    __syncthreads();
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    s = warp_sum((a * a + b * b) + (c * c + d * d));
    if (lane == 0) red[warp] = s;
    __syncthreads();
    const float sd = sqrtf((((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) * (1.0f / 1023.0f)) + 1e-7f;
    {
        const int y = tid >> 3, x0 = (tid & 7) * 4;
        xs[y + 1][x0 + 1] = a / sd; xs[y + 1][x0 + 2] = b / sd; xs[y + 1][x0 + 3] = c / sd; xs[y + 1][x0 + 4] = d / sd;
    }
    __syncthreads();
    unsigned char* dst = reinterpret_cast<unsigned char*>(out) + (size_t)p * (32 * kL1R * EB);
    for (int i = 0; i < 4; ++i) {
        const int px = tid + 256 * i, y = px >> 5, xx = px & 31;
        float acc[32];
#pragma unroll
        for (int cch = 0; cch < 32; ++cch) acc[cch] = ssh[cch];
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            const float xv = xs[y + t / 3][xx + t % 3];
#pragma unroll
            for (int cch = 0; cch < 32; ++cch) acc[cch] = fmaf(xv, wsm[t][cch], acc[cch]);
        }
        const int row = (y + 1) * 34 + xx + 1;
#pragma unroll
        for (int cch = 0; cch < 32; ++cch) acc[cch] = fmaxf(acc[cch], 0.f);
        if (EB == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + ((size_t)j * kL1R + row) * 16) = hn_pack8(acc + 8 * j);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(dst + ((size_t)j * kL1R + row) * 16) =
                    to_tf32(make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
        }
    }
}

// ------------------------------------------------------------------------------------------ final layer
// [128 patches x 8192] x [8192 x 128] on tcgen05, split-K over kFinalSplit CTAs per patch tile.  Both operands arrive as
// pre-laid-out 16 KB blocks (A written by conv6's epilogue, B packed once), one bulk copy each per K block.
template <int EB>
__global__ void __launch_bounds__(kHnThreads, 1)
hn_tc_final_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ partial, int n) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t kBlk = kFinalBlockFloats * (uint32_t)EB;     // 16 KB (tf32) / 8 KB (fp16)
    constexpr int kPer = kFinalBlocks / kFinalSplit;                // K blocks per CTA
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kHnSlots * kBlk);
    uint64_t* full = bars;
    uint64_t* empty = bars + kHnSlots;
    uint64_t* done = bars + 2 * kHnSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kHnSlots + 1);
    const int tid = threadIdx.x, tile = blockIdx.x, split = blockIdx.y;
    const bool w0 = warp0_uniform();
    if (tid < 32) tmem_alloc(tmem_slot, 128);
    if (tid == 0) {
        for (int i = 0; i < kHnSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *tmem_slot;
    if (w0 && elect_one()) {
        const unsigned char* ga = reinterpret_cast<const unsigned char*>(a) + ((size_t)tile * kFinalBlocks + (size_t)split * kPer) * kBlk;
        const unsigned char* gb = reinterpret_cast<const unsigned char*>(b) + (size_t)split * kPer * kBlk;
        const uint32_t s_addr = smem_u32(smem);
        constexpr uint32_t idesc = EB == 2 ? hn_idesc_f16(128, 128) : make_idesc_tf32(128, 128);
        constexpr int KSTEP = 32 / EB;
        int produced = 0;
        for (int blk = 0; blk < kPer; ++blk) {
            while (produced < kPer && produced < blk + kHnSlots) {
                const int slot = produced % kHnSlots, use = produced / kHnSlots;
                if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
                mbar_expect_tx(&full[slot], 2 * kBlk);
                bulk_g2s(smem + (size_t)slot * 2 * kBlk, ga + (size_t)produced * kBlk, kBlk, &full[slot]);
                bulk_g2s(smem + (size_t)slot * 2 * kBlk + kBlk, gb + (size_t)produced * kBlk, kBlk, &full[slot]);
                ++produced;
            }
            const int slot = blk % kHnSlots;
            mbar_wait(&full[slot], (blk / kHnSlots) & 1);
            fence_after_sync();
            const uint64_t a_desc = hn_desc(s_addr + slot * 2 * kBlk, 128u * 16u), b_desc = hn_desc(s_addr + slot * 2 * kBlk + kBlk, 128u * 16u);
#pragma unroll
            for (int k8 = 0; k8 < kFinalKB / KSTEP; ++k8) {
                if (EB == 2) hn_mma_f16(tm, a_desc + ((k8 * 2u * 128u * 16u) >> 4), b_desc + ((k8 * 2u * 128u * 16u) >> 4), idesc, !(blk == 0 && k8 == 0));
                else mma_tf32(tm, a_desc + ((k8 * 2u * 128u * 16u) >> 4), b_desc + ((k8 * 2u * 128u * 16u) >> 4), idesc, !(blk == 0 && k8 == 0));
            }
            commit(&empty[slot]);
        }
        commit(done);
    }
    if (tid == 0) mbar_wait(done, 0);
    __syncthreads();
    fence_after_sync();
    const int p = tile * 128 + tid;
    float* dst = partial + ((size_t)split * n + p) * 128;
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)(tid & ~31) << 16) + c0, v);
        tmem_ld_wait();
        if (p < n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + c0 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 128);
}

// split-K partial sums -> + BatchNorm shift -> L2 normalisation (hardnet_pytorch.py:7-15).  One warp per patch.
__global__ void hn_tc_finish_kernel(const float* __restrict__ partial, const float* __restrict__ shift, int n, float* __restrict__ desc) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n) return;
    float4 v = __ldg(reinterpret_cast<const float4*>(shift) + lane);
#pragma unroll
    for (int s = 0; s < kFinalSplit; ++s) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)s * n + p) * 128) + lane);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    const float nrm = sqrtf(warp_sum((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w)) + 1e-10f);
    reinterpret_cast<float4*>(desc + (size_t)p * 128)[lane] = make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
}

// ------------------------------------------------------------------------------------------ weights
struct HnTcLayer { int cin, cout, kb; };
static const HnTcLayer kHnTc[5] = {{32, 32, 32}, {32, 64, 32}, {64, 64, 64}, {64, 128, 16}, {128, 128, 32}};   // conv2 .. conv6

struct HnTcBlob {
    size_t first_w, first_shift;     // conv1: [9][32] scaled, [32]
    size_t w[5], shift[5];           // conv2..6: NBLK blocks of [cout x kb] chunk-major, [cout]
    size_t final_w, final_shift;     // final layer: kFinalBlocks blocks of [128 x 32] chunk-major, [128]
    size_t w16[5], final_w16;        // the same blocks as fp16 (8 per chunk) for the fp16-operand path
    size_t floats;
};
static HnTcBlob hn_tc_layout() {
    HnTcBlob b;
    size_t off = 0;
    auto take = [&](size_t nfl) { size_t o = off; off = (off + nfl + 63) / 64 * 64; return o; };
    b.first_w = take(288); b.first_shift = take(32);
    for (int l = 0; l < 5; ++l) {
        b.w[l] = take((size_t)9 * kHnTc[l].cin * kHnTc[l].cout);
        b.shift[l] = take(kHnTc[l].cout);
    }
    b.final_w = take((size_t)8192 * 128); b.final_shift = take(128);
    for (int l = 0; l < 5; ++l) b.w16[l] = take((size_t)9 * kHnTc[l].cin * kHnTc[l].cout / 2);
    b.final_w16 = take((size_t)8192 * 128 / 2);
    b.floats = off;
    return b;
}
size_t hn_tc_blob_floats() { return hn_tc_layout().floats; }

// wT [cin][9][cout] (fp32 path layout) * scale[cout] -> blocks (tap, K part) of [cout rows][kb] chunk-major, tf32
__global__ void hn_tc_pack_kernel(const float* __restrict__ wT, const float* __restrict__ scale, int cin, int cout, int kb,
                                  float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * cin * cout) return;
    const int n = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
    const int kp = ci / kb, kk = ci - kp * kb;
    const float v = wT[((size_t)ci * 9 + tap) * cout + n] * scale[n];
    dst[(size_t)(tap * (cin / kb) + kp) * cout * kb + (size_t)(kk >> 2) * cout * 4 + n * 4 + (kk & 3)] = to_tf32_exact(v);
}
// final layer: wT [cin = 128][64 positions][cout = 128] * scale -> K blocks (position, 32 channels) of [128 rows][32], tf32
__global__ void hn_tc_pack_final_kernel(const float* __restrict__ wT, const float* __restrict__ scale, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 8192 * 128) return;
    const int n = i % 128, c = (i / 128) % 128, pos = i / (128 * 128);
    const int blk = pos * 4 + c / 32, kk = c % 32;
    dst[(size_t)blk * kFinalBlockFloats + (size_t)(kk >> 2) * 512 + n * 4 + (kk & 3)] = to_tf32_exact(wT[((size_t)c * 64 + pos) * 128 + n] * scale[n]);
}
// fp16 versions of the two kernels above (8 elements per 16-byte chunk)
__global__ void hn_tc_pack16_kernel(const float* __restrict__ wT, const float* __restrict__ scale, int cin, int cout, int kb,
                                    __half* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * cin * cout) return;
    const int n = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
    const int kp = ci / kb, kk = ci - kp * kb;
    const float v = wT[((size_t)ci * 9 + tap) * cout + n] * scale[n];
    dst[(size_t)(tap * (cin / kb) + kp) * cout * kb + (size_t)(kk >> 3) * cout * 8 + n * 8 + (kk & 7)] = __float2half_rn(v);
}
__global__ void hn_tc_pack16_final_kernel(const float* __restrict__ wT, const float* __restrict__ scale, __half* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 8192 * 128) return;
    const int n = i % 128, c = (i / 128) % 128, pos = i / (128 * 128);
    const int blk = pos * 4 + c / 32, kk = c % 32;
    dst[(size_t)blk * kFinalBlockFloats + (size_t)(kk >> 3) * 1024 + n * 8 + (kk & 7)] = __float2half_rn(wT[((size_t)c * 64 + pos) * 128 + n] * scale[n]);
}
__global__ void hn_tc_pack_first_kernel(const float* __restrict__ wT, const float* __restrict__ scale, float* __restrict__ dst) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < 288) dst[i] = wT[i] * scale[i % 32];
}
__global__ void hn_tc_copy_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < n) dst[i] = src[i];
}

int hn_tc_pack_weights(const HnW& w, float* blob, cudaStream_t st) {
    const HnTcBlob L = hn_tc_layout();
    BALF_CUDA_OK(cudaMemsetAsync(blob, 0, L.floats * sizeof(float), st));
    hn_tc_pack_first_kernel<<<2, 256, 0, st>>>(w.w[0], w.scale[0], blob + L.first_w);
    hn_tc_copy_kernel<<<1, 128, 0, st>>>(w.shift[0], 32, blob + L.first_shift);
    for (int l = 0; l < 5; ++l) {
        const HnTcLayer& T = kHnTc[l];
        hn_tc_pack_kernel<<<cdiv(9 * T.cin * T.cout, 256), 256, 0, st>>>(w.w[l + 1], w.scale[l + 1], T.cin, T.cout, T.kb, blob + L.w[l]);
        hn_tc_copy_kernel<<<1, 128, 0, st>>>(w.shift[l + 1], T.cout, blob + L.shift[l]);
        hn_tc_pack16_kernel<<<cdiv(9 * T.cin * T.cout, 256), 256, 0, st>>>(w.w[l + 1], w.scale[l + 1], T.cin, T.cout, T.kb,
                                                                           reinterpret_cast<__half*>(blob + L.w16[l]));
    }
    hn_tc_pack16_final_kernel<<<cdiv(8192 * 128, 256), 256, 0, st>>>(w.w[6], w.scale[6], reinterpret_cast<__half*>(blob + L.final_w16));
    hn_tc_pack_final_kernel<<<cdiv(8192 * 128, 256), 256, 0, st>>>(w.w[6], w.scale[6], blob + L.final_w);
    hn_tc_copy_kernel<<<1, 128, 0, st>>>(w.shift[6], 128, blob + L.final_shift);
    BALF_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------ host
using G2 = HnGeom<32, 32, 1, 32, 1, 32>;
using G3 = HnGeom<32, 64, 2, 16, 1, 32>;
using G4 = HnGeom<64, 64, 1, 16, 2, 64>;
using G5 = HnGeom<64, 128, 2, 8, 2, 16>;
using G6 = HnGeom<128, 128, 1, 8, 2, 32>;
static_assert(G2::R == 1160 && G3::R == 1160 && G4::R == 328 && G5::R == 384 && G6::R == 152, "layer image sizes");

constexpr int kHnTcChunk = 4096;     // patches per internal pass (bounds the workspace: 0.57 MB per patch)
struct HnTcWs { float* a[7]; };      // inputs of conv2 .. conv6, the final layer's A operand (whole 128-patch tiles), split-K partials
static size_t hn_tc_ws_layout(int n, void* base, HnTcWs* ws) {
    const size_t ntile = (size_t)cdiv(n, 128) * 128;
    const size_t floats[7] = {(size_t)32 * G2::R * n, (size_t)32 * G3::R * n, (size_t)64 * G4::R * n, (size_t)64 * G5::R * n,
                              (size_t)128 * G6::R * n, 8192 * ntile, (size_t)kFinalSplit * 128 * n};
    size_t off = 0;
    for (int i = 0; i < 7; ++i) {
        if (ws) ws->a[i] = reinterpret_cast<float*>(static_cast<char*>(base) + off);
        off = align_up(off + floats[i] * sizeof(float), 256);
    }
    return off;
}
size_t hn_tc_workspace_bytes(int n_patches) { return hn_tc_ws_layout(n_patches < kHnTcChunk ? n_patches : kHnTcChunk, nullptr, nullptr); }

static int hn_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int CIN, int COUT, int STRIDE, int HOUT, int G, int KB, int NEXT, int RN, int EB>
static int hn_tc_launch(const char* name, const float* in, const float* wblk, const float* shift, float* out, int n, cudaStream_t st) {
    using Ge = HnGeom<CIN, COUT, STRIDE, HOUT, G, KB, EB>;
    auto kernel = hn_tc_conv_kernel<CIN, COUT, STRIDE, HOUT, G, KB, NEXT, RN, EB>;
    BALF_REQUIRE(Ge::smem <= 227 * 1024, "internal: %s needs %zu bytes of shared memory", name, Ge::smem);
    BALF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Ge::smem));
    const int ngroups = cdiv(n, G);
    const int grid = ngroups < hn_num_sms() ? ngroups : hn_num_sms();
    {
        ProfScope p(name, st);
        kernel<<<grid, kHnConvThreads, Ge::smem, st>>>(in, wblk, shift, out, n);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

// EB = 4: tf32 operands (precision 1), EB = 2: fp16 operands (precision 2) -- same 11-bit significand, half the bytes of every
// activation image and weight block, half the MMAs (K = 16 per instruction), and room for twice the patches per CTA
template <int EB>
static int hn_tc_forward_t(const float* blob, const float* patches, int n_patches, float* desc, void* workspace, cudaStream_t st) {
    const HnTcBlob L = hn_tc_layout();
    // patches per CTA: bounded by the 512 TMEM columns (G x tiles x COUT) and, at tf32, by shared memory
    constexpr int G3_ = EB == 2 ? 2 : 1, G56 = EB == 2 ? 4 : 2;
    const float* wl[5];
    for (int l = 0; l < 5; ++l) wl[l] = blob + (EB == 2 ? L.w16[l] : L.w[l]);
    for (int p0 = 0; p0 < n_patches; p0 += kHnTcChunk) {
        const int n = n_patches - p0 < kHnTcChunk ? n_patches - p0 : kHnTcChunk;
        HnTcWs ws;
        hn_tc_ws_layout(n_patches < kHnTcChunk ? n_patches : kHnTcChunk, workspace, &ws);
        {
            ProfScope p("hn_tc_conv1", st);
            hn_tc_first_kernel<EB><<<n, 256, 0, st>>>(patches + (size_t)p0 * 1024, blob + L.first_w, blob + L.first_shift, ws.a[0]);
        }
        BALF_COUNT_LAUNCH(1);
        BALF_LAUNCH_OK();
        if (int e = hn_tc_launch<32, 32, 1, 32, 1, 32, 1, G3::R, EB>("hn_tc_conv2", ws.a[0], wl[0], blob + L.shift[0], ws.a[1], n, st)) return e;
        if (int e = hn_tc_launch<32, 64, 2, 16, G3_, 32, 0, G4::R, EB>("hn_tc_conv3", ws.a[1], wl[1], blob + L.shift[1], ws.a[2], n, st)) return e;
        if (int e = hn_tc_launch<64, 64, 1, 16, 2, 64, 1, G5::R, EB>("hn_tc_conv4", ws.a[2], wl[2], blob + L.shift[2], ws.a[3], n, st)) return e;
        if (int e = hn_tc_launch<64, 128, 2, 8, G56, 16, 0, G6::R, EB>("hn_tc_conv5", ws.a[3], wl[3], blob + L.shift[3], ws.a[4], n, st)) return e;
        if (int e = hn_tc_launch<128, 128, 1, 8, G56, 32, 2, 0, EB>("hn_tc_conv6", ws.a[4], wl[4], blob + L.shift[4], ws.a[5], n, st)) return e;
        {
            constexpr size_t smem = 2 * kHnSlots * kFinalBlockFloats * EB + 128;
            BALF_CUDA_OK(cudaFuncSetAttribute(hn_tc_final_kernel<EB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ProfScope p("hn_tc_final", st);
            hn_tc_final_kernel<EB><<<dim3(cdiv(n, 128), kFinalSplit), kHnThreads, smem, st>>>(ws.a[5], blob + (EB == 2 ? L.final_w16 : L.final_w), ws.a[6], n);
            hn_tc_finish_kernel<<<cdiv(n, 8), 256, 0, st>>>(ws.a[6], blob + L.final_shift, n, desc + (size_t)p0 * 128);
        }
        BALF_COUNT_LAUNCH(2);
        BALF_LAUNCH_OK();
    }
    return 0;
}
int hn_tc_forward(const HnW& /*w*/, const float* blob, const float* patches, int n_patches, float* desc, void* workspace, cudaStream_t st,
                  int fp16_operands) {
    return fp16_operands ? hn_tc_forward_t<2>(blob, patches, n_patches, desc, workspace, st)
                         : hn_tc_forward_t<4>(blob, patches, n_patches, desc, workspace, st);
}

}  // namespace balf
