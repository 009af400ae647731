// BALF detector forward on the sm_100a tensor cores: precision 1 ("tf32").
//
// Same semantics and kernel decomposition as the fp32 path in detector.cu (reference:
// balf/model/mlp_ma_decoder.py:201-244 Down, :119-149 multi-axis gMLP, :25-117 grid / block gMLP,
// :151-199 channel attention; balf/model/decoder.py:16-30 head), but every Linear layer and both
// token-mixing products are tcgen05.mma (kind::tf32, fp32 accumulate in TMEM) and the whole chain of
// a tile stays on chip:
//
//   shared memory (A operand, chunk-major) --tcgen05.mma--> TMEM accumulator --tcgen05.ld--> registers
//        ^                                                                          |
//        +---- bias / activation / LayerNorm / gating, one thread per pixel row <---+
//
// A tile is 128 pixels = 2 "units" of 64 tokens (grid branch: the 64 cells of one in-cell offset;
// block branch: one 8x8 block; merge / head: 64 consecutive pixels).  One thread owns one pixel row:
// TMEM lane == thread, so LayerNorm, GELU, softmax and the gating multiply are thread-local.
// The 64x64 token mixing runs as two M=64 MMAs (A = mixing matrix, B = the unit's activations stored
// [channel][token]); their accumulators interleave in the two 16-lane halves of every 32-lane TMEM
// quadrant, which fixes the lane <-> pixel mapping of the branch kernels:
//        unit g = (lane % 32) / 16,   token = (lane / 32) * 16 + lane % 16.
// Weights are pre-packed into the exact shared-memory image of the B operands and brought in by TMA
// bulk copies (cp.async.bulk + mbarrier) -- once per CTA when the whole set fits next to the operand
// region (level 1), otherwise through a ring of 32 KB slots that runs ahead of the MMAs.
#include "detector.cuh"
#include "umma.cuh"

namespace balf {
using namespace umma;

constexpr int TM = 128;                 // pixels per tile == threads per CTA
constexpr int kSlotBytes = 32768;
constexpr int kNSlot = 2;
constexpr int kMaxGemm = 6;

struct TcGemm {
    uint32_t goff;        // float offset of the first block inside the tc blob
    uint16_t nblk;        // K blocks
    uint16_t rows;        // rows of the packed operand (N of a Linear layer, 64 for a mixing matrix)
    uint16_t kb;          // K columns per block
    uint16_t pad;
};
struct TcPlan {
    const float* base;
    TcGemm g[kMaxGemm];
    int ngemm;
    int resident;         // all blocks stay in shared memory for the life of the CTA
    uint32_t bytes;       // total bytes of all blocks (resident footprint)
};

enum { BG_CONV0 = 0, BG_PD1, BG_D1A, BG_D1B, BG_WM, BG_D2, BG_COUNT };
enum { MG_CONV0 = 0, MG_PD2A, MG_PD2B, MG_RC1, MG_RC2, MG_COUNT };
enum { HG_C2 = 0, HG_DENSE, HG_COUNT };
constexpr int kHeadN = 80;              // 65 logits padded to a legal UMMA N (multiple of 16)

__host__ __device__ constexpr int tc_kin(int cin) { return cin < 8 ? 8 : cin; }
// K columns per streamed block: the largest power-of-two divisor of K (>= 8) whose block fits a slot
__host__ __device__ constexpr int tc_kb(int rows, int K) {
    int kb = K;
    while (kb > 8 && (kb * rows * 4 > kSlotBytes || K % kb != 0)) kb /= 2;
    return kb;
}
__host__ __device__ constexpr int tc_cols(int need) { return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }

// ------------------------------------------------------------------------------------------ weight ring (thread 0)
struct Ring {
    uint32_t wsm;          // shared address of the weight area
    uint64_t* full;        // [kNSlot]
    uint64_t* empty;       // [kNSlot]
    uint32_t pg, pb;       // producer cursor (gemm, block)
    uint32_t pcnt, ccnt;   // blocks loaded / consumed so far
    uint32_t to_load;      // blocks still to be requested over the life of the CTA
};

__device__ __forceinline__ void ring_load_one(Ring& r, const TcPlan& p) {
    const TcGemm& g = p.g[r.pg];
    const uint32_t bytes = (uint32_t)g.rows * g.kb * 4u;
    const uint32_t slot = r.pcnt % kNSlot, use = r.pcnt / kNSlot;
    if (use > 0) mbar_wait(&r.empty[slot], (use - 1) & 1);
    mbar_expect_tx(&r.full[slot], bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(r.wsm + slot * kSlotBytes), "l"(p.base + g.goff + (size_t)r.pb * g.rows * g.kb), "r"(bytes),
                    "r"(smem_u32(&r.full[slot])) : "memory");
    ++r.pcnt;
    --r.to_load;
    if (++r.pb == g.nblk) { r.pb = 0; if (++r.pg == (uint32_t)p.ngemm) r.pg = 0; }
}
__device__ __forceinline__ void ring_top_up(Ring& r, const TcPlan& p) {
    while (r.to_load > 0 && r.pcnt < r.ccnt + kNSlot) ring_load_one(r, p);
}
// resident mode: every block of every gemm, once
__device__ __forceinline__ void ring_load_all(Ring& r, const TcPlan& p) {
    mbar_expect_tx(&r.full[0], p.bytes);
    uint32_t off = 0;
    for (int gi = 0; gi < p.ngemm; ++gi) {
        const uint32_t bytes = (uint32_t)p.g[gi].rows * p.g[gi].kb * 4u * p.g[gi].nblk;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(r.wsm + off), "l"(p.base + p.g[gi].goff), "r"(bytes), "r"(smem_u32(&r.full[0])) : "memory");
        off += bytes;
    }
}
__device__ __forceinline__ uint32_t resident_off(const TcPlan& p, int gi) {
    uint32_t off = 0;
    for (int i = 0; i < gi; ++i) off += (uint32_t)p.g[i].rows * p.g[i].kb * 4u * p.g[i].nblk;
    return off;
}

// Thread 0: D[128 x N] (+)= A[128 x K] * W^T.  A: chunk-major at shared address a_addr with a_rows
// physical rows; W: gemm gi of the plan (rows == N).  `first` = overwrite the accumulator.
__device__ __forceinline__ void issue_linear(Ring& r, const TcPlan& p, int gi, uint32_t a_addr, uint32_t a_rows,
                                             uint32_t d_tmem, bool first) {
    const TcGemm& g = p.g[gi];
    const uint32_t idesc = make_idesc_tf32(128, g.rows);
    const uint32_t a_lbo = a_rows * 16u, b_lbo = (uint32_t)g.rows * 16u;
    const uint32_t res_base = p.resident ? r.wsm + resident_off(p, gi) : 0u;
    for (uint32_t b = 0; b < g.nblk; ++b) {
        uint32_t w_addr, slot = 0;
        if (p.resident) {
            w_addr = res_base + b * (uint32_t)g.rows * g.kb * 4u;
        } else {
            ring_top_up(r, p);
            slot = r.ccnt % kNSlot;
            mbar_wait(&r.full[slot], (r.ccnt / kNSlot) & 1);
            w_addr = r.wsm + slot * kSlotBytes;
        }
        fence_after_sync();
        for (uint32_t k8 = 0; k8 < (uint32_t)g.kb / 8u; ++k8) {
            const uint32_t kchunk = (b * g.kb) / 4u + k8 * 2u;
            mma_tf32(d_tmem, make_desc(a_addr + kchunk * a_lbo, a_lbo, 128), make_desc(w_addr + k8 * 2u * b_lbo, b_lbo, 128),
                     idesc, !(first && b == 0 && k8 == 0));
        }
        if (!p.resident) { commit(&r.empty[slot]); ++r.ccnt; }
    }
}

// Thread 0: token mixing of both units.  A = mixing matrix (gemm gi, 64 x 64, one block); B = unit u's
// activations [C rows][64 tokens] chunk-major with cp physical rows at y_addr + u * y_stride.
__device__ __forceinline__ void issue_mix(Ring& r, const TcPlan& p, int gi, uint32_t y_addr, uint32_t y_stride, uint32_t cp,
                                          int C, uint32_t d_tmem) {
    uint32_t w_addr, slot = 0;
    if (p.resident) {
        w_addr = r.wsm + resident_off(p, gi);
    } else {
        ring_top_up(r, p);
        slot = r.ccnt % kNSlot;
        mbar_wait(&r.full[slot], (r.ccnt / kNSlot) & 1);
        w_addr = r.wsm + slot * kSlotBytes;
    }
    fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(64, C);
    const uint32_t b_lbo = cp * 16u;
    for (uint32_t u = 0; u < 2; ++u)
        for (uint32_t k8 = 0; k8 < 8; ++k8)
            mma_tf32(d_tmem + ((u * 16u) << 16), make_desc(w_addr + k8 * 2u * 1024u, 1024, 128),
                     make_desc(y_addr + u * y_stride + k8 * 2u * b_lbo, b_lbo, 128), idesc, k8 > 0);
    if (!p.resident) { commit(&r.empty[slot]); ++r.ccnt; }
}

// ------------------------------------------------------------------------------------------ epilogue pieces
template <int ACT> __device__ __forceinline__ float tc_act(float v) {
    if (ACT == 1) return fmaxf(v, 0.0f);
    if (ACT == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    if (ACT == 3) return v > 0.0f ? v : 0.2f * v;
    return v;
}

// v = act(acc + bias) for C columns starting at taddr; optionally parked back in TMEM; returns sum / sum of squares
template <int C, int ACT, bool PARK>
__device__ __forceinline__ void epi_act(uint32_t taddr, const float* __restrict__ bias, float& sum, float& sq) {
    sum = 0.f; sq = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v[i] = tc_act<ACT>(v[i] + __ldg(bias + c0 + i));
            sum += v[i];
            sq = fmaf(v[i], v[i], sq);
        }
        if (PARK) tmem_st32(taddr + c0, v);
    }
    if (PARK) tmem_st_wait();
}

__device__ __forceinline__ void ln_stats(float sum, float sq, int C, float& mean, float& rstd) {
    mean = sum / (float)C;
    const float var = fmaxf(sq / (float)C - mean * mean, 0.f);
    rstd = 1.0f / sqrtf(var + 1e-5f);
}

// LayerNorm of the parked row -> A operand (chunk-major, 128 rows) at `dst`, row `row`
template <int C>
__device__ __forceinline__ void epi_ln_to_a(uint32_t taddr, float mean, float rstd, const float* __restrict__ gam,
                                            const float* __restrict__ bet, float* dst, int row) {
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = (v[4 * j + 0] - mean) * rstd * __ldg(gam + c0 + 4 * j + 0) + __ldg(bet + c0 + 4 * j + 0);
            o.y = (v[4 * j + 1] - mean) * rstd * __ldg(gam + c0 + 4 * j + 1) + __ldg(bet + c0 + 4 * j + 1);
            o.z = (v[4 * j + 2] - mean) * rstd * __ldg(gam + c0 + 4 * j + 2) + __ldg(bet + c0 + 4 * j + 2);
            o.w = (v[4 * j + 3] - mean) * rstd * __ldg(gam + c0 + 4 * j + 3) + __ldg(bet + c0 + 4 * j + 3);
            *reinterpret_cast<float4*>(dst + ((size_t)(c0 / 4 + j) * TM + row) * 4) = to_tf32(o);
        }
    }
}

// ------------------------------------------------------------------------------------------ shared memory carve-up
struct TcShared {
    float* region;         // operand region
    uint32_t wsm;          // weight area (shared address)
    uint64_t* full;
    uint64_t* empty;
    uint64_t* done;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ TcShared carve(unsigned char* smem, uint32_t region_bytes, uint32_t weight_bytes) {
    TcShared s;
    s.region = reinterpret_cast<float*>(smem);
    s.wsm = smem_u32(smem + region_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + region_bytes + weight_bytes);
    s.full = bars;
    s.empty = bars + kNSlot;
    s.done = bars + 2 * kNSlot;
    s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNSlot + 1);
    return s;
}
__host__ __device__ constexpr uint32_t tc_weight_bytes(bool resident, uint32_t total) { return resident ? (total + 127u) / 128u * 128u : kNSlot * kSlotBytes; }
constexpr uint32_t kTcTail = 128;      // barriers + TMEM slot

__device__ __forceinline__ void tc_prologue(const TcShared& s, uint32_t ncols, Ring& ring, const TcPlan& plan, uint32_t my_tiles) {
    if (threadIdx.x < 32) tmem_alloc(s.tmem_slot, ncols);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNSlot; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        mbar_init(s.done, 1);
        mbar_fence_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    ring.wsm = s.wsm; ring.full = s.full; ring.empty = s.empty;
    ring.pg = ring.pb = ring.pcnt = ring.ccnt = 0;
    uint32_t nb = 0;
    for (int i = 0; i < plan.ngemm; ++i) nb += plan.g[i].nblk;
    ring.to_load = nb * my_tiles;
    if (threadIdx.x == 0) {
        if (plan.resident) { ring_load_all(ring, plan); mbar_wait(&s.full[0], 0); }
        else ring_top_up(ring, plan);
    }
}
__device__ __forceinline__ void tc_epilogue_done(const TcShared& s, uint32_t tm, uint32_t ncols) {
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, ncols);
}
// all threads: operand region written -> visible to the tensor core, TMEM reads retired
__device__ __forceinline__ void sync_for_mma() {
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}
__device__ __forceinline__ void wait_done(uint64_t* done, uint32_t& phase) {
    mbar_wait(done, phase & 1);
    ++phase;
    fence_after_sync();
}

struct UnitGeom {
    int h, w, fh, fw;       // level size, grid-cell extent
    int upi;                // units (64 tokens) per image
    int total_units;        // over the batch chunk
};

// pixel index (inside its image) of token `tok` of unit `u`
template <int KIND>   // 0 grid, 1 block, 2 linear
__device__ __forceinline__ int unit_pixel(const UnitGeom& g, int u, int tok) {
    if (KIND == 0) { const int fy = u / g.fw, fx = u - fy * g.fw; return ((tok >> 3) * g.fh + fy) * g.w + (tok & 7) * g.fw + fx; }
    if (KIND == 1) { const int bw = g.w >> 3, by = u / bw, bx = u - by * bw; return (by * 8 + (tok >> 3)) * g.w + bx * 8 + (tok & 7); }
    return u * 64 + tok;
}

// this thread's pixel row of the level input -> A operand (chunk-major, 128 rows), K padded to KIN
template <int CIN>
__device__ __forceinline__ void load_input_row(const float* __restrict__ xin, bool nchw, size_t npix, size_t img, int pix,
                                               bool valid, float* dst, int row) {
    constexpr int KIN = tc_kin(CIN);
    if (CIN < 8) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (valid && c < CIN) ? __ldg(xin + (img * CIN + c) * npix + pix) : 0.f;
        *reinterpret_cast<float4*>(dst + ((size_t)0 * TM + row) * 4) = to_tf32(make_float4(v[0], v[1], v[2], v[3]));
        *reinterpret_cast<float4*>(dst + ((size_t)1 * TM + row) * 4) = to_tf32(make_float4(v[4], v[5], v[6], v[7]));
    } else {
        (void)nchw;
        const float4* src = reinterpret_cast<const float4*>(xin + (img * npix + pix) * CIN);
#pragma unroll 4
        for (int j = 0; j < KIN / 4; ++j)
            *reinterpret_cast<float4*>(dst + ((size_t)j * TM + row) * 4) = valid ? to_tf32(__ldg(src + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------ branch kernel
template <int C> struct BranchCfg {
    static constexpr int CP = C + 1;                               // padded rows of the [channel][token] operand
    static constexpr uint32_t y_stride = (CP * 64 + 16) * 4;       // bytes between the two units' operands
    static constexpr uint32_t region = (2 * y_stride > (uint32_t)TM * C * 4 ? 2 * y_stride : (uint32_t)TM * C * 4);
    static constexpr bool park_u = C <= 128;                       // u stays in TMEM (else it round-trips through `out`)
    static constexpr int col_u = 0;
    static constexpr int col_y = park_u ? C : 0;
    static constexpr int ncols = tc_cols(col_y + 2 * C);
};

template <int CIN, int C, int BR>
__global__ void __launch_bounds__(TM, 1) tc_branch_kernel(const float* __restrict__ xin, int in_nchw, DownW w, TcPlan plan,
                                                          UnitGeom geo, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using Cfg = BranchCfg<C>;
    const TcShared s = carve(smem, Cfg::region, tc_weight_bytes(plan.resident, plan.bytes));
    const int tid = threadIdx.x;
    const int ntiles = (geo.total_units + 1) / 2;
    const uint32_t my_tiles = blockIdx.x < (unsigned)ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    Ring ring;
    tc_prologue(s, Cfg::ncols, ring, plan, my_tiles);
    const uint32_t tm = *s.tmem_slot;
    const uint32_t lane_base = tm + ((uint32_t)(tid & ~31) << 16);       // this warp's 32-lane window
    const uint32_t region_addr = smem_u32(s.region);
    const int ug = (tid & 31) >> 4, tok = (tid >> 5) * 16 + (tid & 15);  // unit / token of this lane
    const DownW::Branch& br = w.br[BR];
    const float mix_bias = __ldg(br.gd_b + tok);
    const size_t npix = (size_t)geo.h * geo.w;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int unit = 2 * t + ug;
        const bool valid = unit < geo.total_units;
        const int img = valid ? unit / geo.upi : 0, u = valid ? unit - img * geo.upi : 0;
        const int pix = unit_pixel<BR>(geo, u, tok);
        // ---- x -> conv.0 -> ReLU -> LayerNorm
        load_input_row<CIN>(xin, in_nchw != 0, npix, (size_t)img, pix, valid, s.region, tid);
        sync_for_mma();
        if (tid == 0) { issue_linear(ring, plan, BG_CONV0, region_addr, TM, tm + Cfg::col_y, true); commit(s.done); }
        wait_done(s.done, phase);
        float sum, sq, mean, rstd;
        epi_act<C, 1, true>(lane_base + Cfg::col_y, w.conv0_b, sum, sq);
        ln_stats(sum, sq, C, mean, rstd);
        epi_ln_to_a<C>(lane_base + Cfg::col_y, mean, rstd, w.pn_w, w.pn_b, s.region, tid);
        sync_for_mma();
        // ---- this branch's half of dense1 -> GELU = u (residual) -> LayerNorm
        if (tid == 0) { issue_linear(ring, plan, BG_PD1, region_addr, TM, tm + Cfg::col_u, true); commit(s.done); }
        wait_done(s.done, phase);
        epi_act<C, 2, true>(lane_base + Cfg::col_u, w.pd1_b + BR * C, sum, sq);
        ln_stats(sum, sq, C, mean, rstd);
        float* orow = out + ((size_t)img * npix + pix) * C;
        if (!Cfg::park_u) {                                        // u round-trips through the output row
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 32) {                   // (tcgen05.ld is warp-collective: never under `valid`)
                float v[32];
                tmem_ld32(lane_base + Cfg::col_u + c0, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(orow + c0 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
        epi_ln_to_a<C>(lane_base + Cfg::col_u, mean, rstd, br.n_w, br.n_b, s.region, tid);
        sync_for_mma();
        // ---- gMLP dense1 (two halves) -> GELU; y1 parked, y2 -> LayerNorm -> [channel][token] operand
        if (tid == 0) {
            issue_linear(ring, plan, BG_D1A, region_addr, TM, tm + Cfg::col_y, true);
            issue_linear(ring, plan, BG_D1B, region_addr, TM, tm + Cfg::col_y + C, true);
            commit(s.done);
        }
        wait_done(s.done, phase);
        float s1, q1;
        epi_act<C, 2, true>(lane_base + Cfg::col_y, br.d1_b, s1, q1);
        epi_act<C, 2, true>(lane_base + Cfg::col_y + C, br.d1_b + C, sum, sq);
        ln_stats(sum, sq, C, mean, rstd);
        {
            float* yt = s.region + (size_t)ug * (Cfg::y_stride / 4) + (size_t)(tok >> 2) * (Cfg::CP * 4) + (tok & 3);
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 32) {
                float v[32];
                tmem_ld32(lane_base + Cfg::col_y + C + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    yt[(size_t)(c0 + i) * 4] = to_tf32((v[i] - mean) * rstd * __ldg(br.gn_w + c0 + i) + __ldg(br.gn_b + c0 + i));
            }
        }
        sync_for_mma();
        // ---- token mixing, gating y1 * (y2' + 1)
        if (tid == 0) { issue_mix(ring, plan, BG_WM, region_addr, Cfg::y_stride, Cfg::CP, C, tm + Cfg::col_y + C); commit(s.done); }
        wait_done(s.done, phase);
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float y1[32], y2[32];
            tmem_ld32(lane_base + Cfg::col_y + c0, y1);
            tmem_ld32(lane_base + Cfg::col_y + C + c0, y2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = y1[4 * j + 0] * (y2[4 * j + 0] + mix_bias + 1.0f);
                o.y = y1[4 * j + 1] * (y2[4 * j + 1] + mix_bias + 1.0f);
                o.z = y1[4 * j + 2] * (y2[4 * j + 2] + mix_bias + 1.0f);
                o.w = y1[4 * j + 3] * (y2[4 * j + 3] + mix_bias + 1.0f);
                *reinterpret_cast<float4*>(s.region + ((size_t)(c0 / 4 + j) * TM + tid) * 4) = to_tf32(o);
            }
        }
        sync_for_mma();
        // ---- dense2 + residual u -> out
        if (tid == 0) { issue_linear(ring, plan, BG_D2, region_addr, TM, tm + Cfg::col_y, true); commit(s.done); }
        wait_done(s.done, phase);
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32], r[32];
            tmem_ld32(lane_base + Cfg::col_y + c0, a);
            if (Cfg::park_u) tmem_ld32(lane_base + Cfg::col_u + c0, r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 res;
                    if (Cfg::park_u) res = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    else res = *reinterpret_cast<const float4*>(orow + c0 + 4 * j);
                    float4 o;
                    o.x = a[4 * j + 0] + __ldg(br.d2_b + c0 + 4 * j + 0) + res.x;
                    o.y = a[4 * j + 1] + __ldg(br.d2_b + c0 + 4 * j + 1) + res.y;
                    o.z = a[4 * j + 2] + __ldg(br.d2_b + c0 + 4 * j + 2) + res.z;
                    o.w = a[4 * j + 3] + __ldg(br.d2_b + c0 + 4 * j + 3) + res.w;
                    *reinterpret_cast<float4*>(orow + c0 + 4 * j) = o;
                }
            }
        }
        // the next tile's input load overwrites the region: every MMA reading it has completed (wait_done)
    }
    tc_epilogue_done(s, tm, Cfg::ncols);
}

// ------------------------------------------------------------------------------------------ merge kernel
template <int C> struct MergeCfg {
    static constexpr uint32_t region = (uint32_t)TM * C * 4;
    static constexpr int col_x0 = 0, col_acc = C;
    static constexpr int ncols = tc_cols(2 * C);
};

template <int CIN, int C>
__global__ void __launch_bounds__(TM, 1) tc_merge_kernel(const float* __restrict__ xin, int in_nchw, DownW w, TcPlan plan,
                                                         UnitGeom geo, const float* __restrict__ uin, const float* __restrict__ vin,
                                                         float* __restrict__ rout, float* __restrict__ qout,
                                                         float* __restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using Cfg = MergeCfg<C>;
    const TcShared s = carve(smem, Cfg::region, tc_weight_bytes(plan.resident, plan.bytes));
    const int tid = threadIdx.x;
    const int ntiles = (geo.total_units + 1) / 2;
    const uint32_t my_tiles = blockIdx.x < (unsigned)ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    Ring ring;
    tc_prologue(s, Cfg::ncols, ring, plan, my_tiles);
    const uint32_t tm = *s.tmem_slot;
    const uint32_t lane_base = tm + ((uint32_t)(tid & ~31) << 16);
    const uint32_t region_addr = smem_u32(s.region);
    const int ug = tid >> 6, tok = tid & 63;
    const size_t npix = (size_t)geo.h * geo.w;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int unit = 2 * t + ug;
        const bool valid = unit < geo.total_units;
        const int img = valid ? unit / geo.upi : 0, u = valid ? unit - img * geo.upi : 0;
        const int pix = u * 64 + tok;
        const size_t row_off = ((size_t)img * npix + pix) * C;
        // ---- x0 = ReLU(conv.0(x)), parked
        load_input_row<CIN>(xin, in_nchw != 0, npix, (size_t)img, pix, valid, s.region, tid);
        sync_for_mma();
        if (tid == 0) { issue_linear(ring, plan, MG_CONV0, region_addr, TM, tm + Cfg::col_x0, true); commit(s.done); }
        wait_done(s.done, phase);
        float sum, sq, mean, rstd;
        epi_act<C, 1, true>(lane_base + Cfg::col_x0, w.conv0_b, sum, sq);
        // ---- dense2([u', v']) accumulated over the two K halves (the region is reloaded in between)
        load_input_row<C>(uin, false, npix, (size_t)img, pix, valid, s.region, tid);
        sync_for_mma();
        if (tid == 0) { issue_linear(ring, plan, MG_PD2A, region_addr, TM, tm + Cfg::col_acc, true); commit(s.done); }
        wait_done(s.done, phase);
        load_input_row<C>(vin, false, npix, (size_t)img, pix, valid, s.region, tid);
        sync_for_mma();
        if (tid == 0) { issue_linear(ring, plan, MG_PD2B, region_addr, TM, tm + Cfg::col_acc, false); commit(s.done); }
        wait_done(s.done, phase);
        // x1 = acc + b + x0 (parked over the accumulator); q = x1 + x0 -> global; LayerNorm(x1) -> region
        sum = 0.f; sq = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32], x0[32];
            tmem_ld32(lane_base + Cfg::col_acc + c0, a);
            tmem_ld32(lane_base + Cfg::col_x0 + c0, x0);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                a[i] = a[i] + __ldg(w.pd2_b + c0 + i) + x0[i];
                sum += a[i];
                sq = fmaf(a[i], a[i], sq);
                x0[i] = a[i] + x0[i];
            }
            tmem_st32(lane_base + Cfg::col_acc + c0, a);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(qout + row_off + c0 + 4 * j) = make_float4(x0[4 * j], x0[4 * j + 1], x0[4 * j + 2], x0[4 * j + 3]);
            }
        }
        tmem_st_wait();
        ln_stats(sum, sq, C, mean, rstd);
        epi_ln_to_a<C>(lane_base + Cfg::col_acc, mean, rstd, w.rn_w, w.rn_b, s.region, tid);
        sync_for_mma();
        // ---- conv1 -> LeakyReLU(0.2)
        if (tid == 0) { issue_linear(ring, plan, MG_RC1, region_addr, TM, tm + Cfg::col_acc, true); commit(s.done); }
        wait_done(s.done, phase);
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(lane_base + Cfg::col_acc + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = tc_act<3>(a[4 * j + 0] + __ldg(w.rc1_b + c0 + 4 * j + 0));
                o.y = tc_act<3>(a[4 * j + 1] + __ldg(w.rc1_b + c0 + 4 * j + 1));
                o.z = tc_act<3>(a[4 * j + 2] + __ldg(w.rc1_b + c0 + 4 * j + 2));
                o.w = tc_act<3>(a[4 * j + 3] + __ldg(w.rc1_b + c0 + 4 * j + 3));
                *reinterpret_cast<float4*>(s.region + ((size_t)(c0 / 4 + j) * TM + tid) * 4) = to_tf32(o);
            }
        }
        sync_for_mma();
        // ---- conv2 = r -> global, and staged in the region for the per-unit channel sums (squeeze)
        if (tid == 0) { issue_linear(ring, plan, MG_RC2, region_addr, TM, tm + Cfg::col_acc, true); commit(s.done); }
        wait_done(s.done, phase);
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(lane_base + Cfg::col_acc + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = a[4 * j + 0] + __ldg(w.rc2_b + c0 + 4 * j + 0);
                o.y = a[4 * j + 1] + __ldg(w.rc2_b + c0 + 4 * j + 1);
                o.z = a[4 * j + 2] + __ldg(w.rc2_b + c0 + 4 * j + 2);
                o.w = a[4 * j + 3] + __ldg(w.rc2_b + c0 + 4 * j + 3);
                *reinterpret_cast<float4*>(s.region + ((size_t)(c0 / 4 + j) * TM + tid) * 4) = o;
                if (valid) *reinterpret_cast<float4*>(rout + row_off + c0 + 4 * j) = o;
            }
        }
        __syncthreads();
        // channel sums of each unit, fixed order (rotated start so that a warp's lanes hit distinct banks)
        for (int i = tid; i < 2 * C; i += TM) {
            const int uu = i / C, c = i - uu * C;
            const int un = 2 * t + uu;
            if (un < geo.total_units) {
                const float* col = s.region + (size_t)(c >> 2) * TM * 4 + (size_t)uu * 64 * 4 + (c & 3);
                float acc = 0.f;
                for (int k = 0; k < 64; ++k) acc += col[(size_t)((k + (c >> 2)) & 63) * 4];
                partial[(size_t)un * C + c] = acc;
            }
        }
        __syncthreads();
    }
    tc_epilogue_done(s, tm, Cfg::ncols);
}

// ------------------------------------------------------------------------------------------ head kernel (last stage)
// t = r * s + q -> conv2 (C -> C) -> ReLU -> dense (C -> 65) -> folded BatchNorm = logits -> softmax ->
// drop the dustbin -> depth-to-space.  One thread = one 8x8 cell.
template <int C>
__global__ void __launch_bounds__(TM, 1) tc_head_kernel(const float* __restrict__ r, const float* __restrict__ q,
                                                        const float* __restrict__ scale, DownW w, HeadW hw, TcPlan plan,
                                                        UnitGeom geo, int cell, float* __restrict__ logits, float* __restrict__ prob) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t region_bytes = (uint32_t)TM * C * 4;
    constexpr int ncols = tc_cols(C + kHeadN);
    const TcShared s = carve(smem, region_bytes, tc_weight_bytes(plan.resident, plan.bytes));
    const int tid = threadIdx.x;
    const int ntiles = (geo.total_units + 1) / 2;
    const uint32_t my_tiles = blockIdx.x < (unsigned)ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    Ring ring;
    tc_prologue(s, ncols, ring, plan, my_tiles);
    const uint32_t tm = *s.tmem_slot;
    const uint32_t lane_base = tm + ((uint32_t)(tid & ~31) << 16);
    const uint32_t region_addr = smem_u32(s.region);
    const int ug = tid >> 6, tok = tid & 63;
    const size_t npix = (size_t)geo.h * geo.w;
    const int nlog = cell * cell + 1;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int unit = 2 * t + ug;
        const bool valid = unit < geo.total_units;
        const int img = valid ? unit / geo.upi : 0, u = valid ? unit - img * geo.upi : 0;
        const int pix = u * 64 + tok;
        const size_t row_off = ((size_t)img * npix + pix) * C;
#pragma unroll 4
        for (int j = 0; j < C / 4; ++j) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4 rv = __ldg(reinterpret_cast<const float4*>(r + row_off) + j);
                const float4 qv = __ldg(reinterpret_cast<const float4*>(q + row_off) + j);
                const float4 sv = __ldg(reinterpret_cast<const float4*>(scale + (size_t)img * C) + j);
                o = make_float4(rv.x * sv.x + qv.x, rv.y * sv.y + qv.y, rv.z * sv.z + qv.z, rv.w * sv.w + qv.w);
            }
            *reinterpret_cast<float4*>(s.region + ((size_t)j * TM + tid) * 4) = to_tf32(o);
        }
        sync_for_mma();
        if (tid == 0) { issue_linear(ring, plan, HG_C2, region_addr, TM, tm, true); commit(s.done); }
        wait_done(s.done, phase);
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(lane_base + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = fmaxf(a[4 * j + 0] + __ldg(w.c2_b + c0 + 4 * j + 0), 0.f);
                o.y = fmaxf(a[4 * j + 1] + __ldg(w.c2_b + c0 + 4 * j + 1), 0.f);
                o.z = fmaxf(a[4 * j + 2] + __ldg(w.c2_b + c0 + 4 * j + 2), 0.f);
                o.w = fmaxf(a[4 * j + 3] + __ldg(w.c2_b + c0 + 4 * j + 3), 0.f);
                *reinterpret_cast<float4*>(s.region + ((size_t)(c0 / 4 + j) * TM + tid) * 4) = to_tf32(o);
            }
        }
        sync_for_mma();
        if (tid == 0) { issue_linear(ring, plan, HG_DENSE, region_addr, TM, tm + C, true); commit(s.done); }
        wait_done(s.done, phase);
        // logits: folded BatchNorm; softmax over nlog channels (3 passes over 65 parked values)
        float mx = kNegInf;
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += 32) {
            if (c0 >= kHeadN) break;
            float a[32];
            tmem_ld32(lane_base + C + c0, a);      // columns beyond kHeadN hold stale data and are masked below
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = c0 + i;
                if (n < nlog) {
                    const float z = (a[i] + __ldg(hw.b + n)) * __ldg(hw.alpha + n) + __ldg(hw.beta + n);
                    a[i] = z;
                    mx = fmaxf(mx, z);
                    if (valid && logits) logits[((size_t)img * nlog + n) * npix + pix] = z;
                } else a[i] = kNegInf;
            }
            tmem_st32(lane_base + C + c0, a);      // 32 wide: columns up to C + 96 are inside the allocation
        }
        tmem_st_wait();
        float den = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += 32) {
            float a[32];
            tmem_ld32(lane_base + C + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) if (c0 + i < nlog) den += expf(a[i] - mx);
        }
        const int cy = pix / geo.w, cx = pix - cy * geo.w;
        const int Wp = geo.w * cell, Hp = geo.h * cell;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float a[32];
            tmem_ld32(lane_base + C + c0, a);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = c0 + i;
                    if (n < nlog - 1) {
                        const int yy = cy * cell + n / cell, xx = cx * cell + n % cell;
                        prob[((size_t)img * Hp + yy) * Wp + xx] = expf(a[i] - mx) / den;
                    }
                }
            }
        }
    }
    tc_epilogue_done(s, tm, ncols);
}

// ------------------------------------------------------------------------------------------ weight packing (tc blob)
// wT [K][ld] (the fp32 path's transposed weight) -> blocks of [rows x kb] chunk-major, rows n0..n0+rows
__global__ void tc_pack_kernel(const float* __restrict__ wT, int ld, int n0, int rows, int k_real, int k_pad, int kb,
                               float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * k_pad) return;
    const int k = i / rows, n = i - k * rows;
    const int b = k / kb, kk = k - b * kb;
    dst[(size_t)b * rows * kb + (size_t)(kk >> 2) * rows * 4 + n * 4 + (kk & 3)] = k < k_real ? to_tf32(wT[(size_t)k * ld + n0 + n]) : 0.f;
}

struct TcPlans {
    TcPlan branch[4][2], merge[4], head;
    size_t floats;
};

static void tc_add(TcPlan& p, int gi, size_t& off, int rows, int K) {
    const int kb = tc_kb(rows, K);
    p.g[gi].goff = (uint32_t)off;
    p.g[gi].nblk = (uint16_t)(K / kb);
    p.g[gi].rows = (uint16_t)rows;
    p.g[gi].kb = (uint16_t)kb;
    p.g[gi].pad = 0;
    off += (size_t)rows * K;
    p.bytes += (uint32_t)rows * K * 4u;
}

static void tc_build_plans(const balf_detector_arch& a, const float* base, TcPlans* out) {
    size_t off = 0;
    TcPlans P;
    for (int l = 0; l < 4; ++l) {
        const int cin = tc_kin(a.dims[l]), c = a.dims[l + 1];
        const bool resident = c <= 32;
        for (int b = 0; b < 2; ++b) {
            TcPlan& p = P.branch[l][b];
            p = TcPlan{};
            p.base = base; p.ngemm = BG_COUNT; p.resident = resident;
            tc_add(p, BG_CONV0, off, c, cin);
            tc_add(p, BG_PD1, off, c, c);
            tc_add(p, BG_D1A, off, c, c);
            tc_add(p, BG_D1B, off, c, c);
            tc_add(p, BG_WM, off, 64, 64);
            tc_add(p, BG_D2, off, c, c);
        }
        TcPlan& m = P.merge[l];
        m = TcPlan{};
        m.base = base; m.ngemm = MG_COUNT; m.resident = resident;
        tc_add(m, MG_CONV0, off, c, cin);
        tc_add(m, MG_PD2A, off, c, c);
        tc_add(m, MG_PD2B, off, c, c);
        tc_add(m, MG_RC1, off, c, c);
        tc_add(m, MG_RC2, off, c, c);
    }
    TcPlan& h = P.head;
    h = TcPlan{};
    h.base = base; h.ngemm = HG_COUNT; h.resident = 0;
    tc_add(h, HG_C2, off, a.dims[4], a.dims[4]);
    tc_add(h, HG_DENSE, off, kHeadN, a.dims[4]);
    P.floats = off;
    if (out) *out = P;
}

size_t tc_blob_floats(const balf_detector_arch& a) {
    TcPlans P;
    tc_build_plans(a, nullptr, &P);
    return P.floats;
}

static void tc_pack_one(const TcPlan& p, int gi, const float* wT, int ld, int n0, int k_real, float* blob, cudaStream_t st) {
    const TcGemm& g = p.g[gi];
    const int k_pad = g.nblk * g.kb;
    tc_pack_kernel<<<cdiv(g.rows * k_pad, 256), 256, 0, st>>>(wT, ld, n0, g.rows, k_real, k_pad, g.kb, blob + g.goff);
}

// fp32-path packed weights (DetW) -> tc blob
int tc_pack_weights(const balf_detector_arch& a, const DetW& w, float* blob, cudaStream_t st) {
    TcPlans P;
    tc_build_plans(a, blob, &P);
    BALF_CUDA_OK(cudaMemsetAsync(blob, 0, P.floats * sizeof(float), st));
    for (int l = 0; l < 4; ++l) {
        const int ci = a.dims[l], c = a.dims[l + 1];
        const DownW& d = w.down[l];
        for (int b = 0; b < 2; ++b) {
            const TcPlan& p = P.branch[l][b];
            const DownW::Branch& r = d.br[b];
            tc_pack_one(p, BG_CONV0, d.conv0_w, c, 0, ci, blob, st);
            tc_pack_one(p, BG_PD1, d.pd1_w, 2 * c, b * c, c, blob, st);
            tc_pack_one(p, BG_D1A, r.d1_w, 2 * c, 0, c, blob, st);
            tc_pack_one(p, BG_D1B, r.d1_w, 2 * c, c, c, blob, st);
            tc_pack_one(p, BG_WM, r.gd_w, 64, 0, 64, blob, st);
            tc_pack_one(p, BG_D2, r.d2_w, c, 0, c, blob, st);
        }
        const TcPlan& m = P.merge[l];
        tc_pack_one(m, MG_CONV0, d.conv0_w, c, 0, ci, blob, st);
        tc_pack_one(m, MG_PD2A, d.pd2_w, c, 0, c, blob, st);
        tc_pack_one(m, MG_PD2B, d.pd2_w + (size_t)c * c, c, 0, c, blob, st);
        tc_pack_one(m, MG_RC1, d.rc1_w, c, 0, c, blob, st);
        tc_pack_one(m, MG_RC2, d.rc2_w, c, 0, c, blob, st);
    }
    tc_pack_one(P.head, HG_C2, w.down[3].c2_w, a.dims[4], 0, a.dims[4], blob, st);
    tc_pack_one(P.head, HG_DENSE, w.head.w, kHeadPad, 0, a.dims[4], blob, st);
    BALF_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------ host: one stage
static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// persistent grid: one CTA per resident slot.  Resident CTAs per SM = min over shared memory (227 KB usable,
// 1 KB reserved per CTA), registers (64 K per SM) and TMEM columns (512 per SM).
template <typename K>
static int tc_launch_cfg(K kernel, size_t smem, int tmem_cols, int ntiles, int* grid) {
    BALF_REQUIRE(smem <= 227 * 1024, "internal: tc kernel needs %zu bytes of shared memory", smem);
    BALF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BALF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    BALF_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * TM;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (regs_per_cta > 0 && 65536 / regs_per_cta < per_sm) per_sm = 65536 / regs_per_cta;
    if (512 / tmem_cols < per_sm) per_sm = 512 / tmem_cols;
    if (per_sm < 1) per_sm = 1;
    const int cap = num_sms() * per_sm;
    *grid = ntiles < cap ? ntiles : cap;
    return 0;
}

template <int CIN, int C>
static int tc_run_level(const float* xin, bool nchw, const DownW& w, const TcPlans& P, int level, int Bc, int h, int wd,
                        float* u, float* v, float* r, float* q, float* partial, cudaStream_t st) {
    UnitGeom g{h, wd, h / 8, wd / 8, h * wd / 64, Bc * (h * wd / 64)};
    const int ntiles = (g.total_units + 1) / 2;
    int grid = 0;
    for (int b = 0; b < 2; ++b) {
        const TcPlan& p = P.branch[level][b];
        const size_t smem = BranchCfg<C>::region + tc_weight_bytes(p.resident, p.bytes) + kTcTail;
        if (b == 0) {
            if (int e = tc_launch_cfg(tc_branch_kernel<CIN, C, 0>, smem, BranchCfg<C>::ncols, ntiles, &grid)) return e;
            ProfScope ps(C == 32 ? "det_branch_grid_c32" : C == 64 ? "det_branch_grid_c64" : C == 128 ? "det_branch_grid_c128" : "det_branch_grid_c256", st);
            tc_branch_kernel<CIN, C, 0><<<grid, TM, smem, st>>>(xin, nchw, w, p, g, u);
        } else {
            if (int e = tc_launch_cfg(tc_branch_kernel<CIN, C, 1>, smem, BranchCfg<C>::ncols, ntiles, &grid)) return e;
            ProfScope ps(C == 32 ? "det_branch_block_c32" : C == 64 ? "det_branch_block_c64" : C == 128 ? "det_branch_block_c128" : "det_branch_block_c256", st);
            tc_branch_kernel<CIN, C, 1><<<grid, TM, smem, st>>>(xin, nchw, w, p, g, v);
        }
    }
    {
        const TcPlan& p = P.merge[level];
        const size_t smem = MergeCfg<C>::region + tc_weight_bytes(p.resident, p.bytes) + kTcTail;
        if (int e = tc_launch_cfg(tc_merge_kernel<CIN, C>, smem, MergeCfg<C>::ncols, ntiles, &grid)) return e;
        ProfScope ps(C == 32 ? "det_merge_c32" : C == 64 ? "det_merge_c64" : C == 128 ? "det_merge_c128" : "det_merge_c256", st);
        tc_merge_kernel<CIN, C><<<grid, TM, smem, st>>>(xin, nchw, w, p, g, u, v, r, q, partial);
    }
    BALF_COUNT_LAUNCH(3);
    BALF_LAUNCH_OK();
    return 0;
}

int tc_run_level_dispatch(int level, const float* xin, bool nchw, const DownW& w, const balf_detector_arch& a, const float* blob,
                          int Bc, int h, int wd, float* u, float* v, float* r, float* q, float* partial, cudaStream_t st) {
    TcPlans P;
    tc_build_plans(a, blob, &P);
    switch (level) {
        case 0: return tc_run_level<3, 32>(xin, nchw, w, P, 0, Bc, h, wd, u, v, r, q, partial, st);
        case 1: return tc_run_level<32, 64>(xin, nchw, w, P, 1, Bc, h, wd, u, v, r, q, partial, st);
        case 2: return tc_run_level<64, 128>(xin, nchw, w, P, 2, Bc, h, wd, u, v, r, q, partial, st);
        default: return tc_run_level<128, 256>(xin, nchw, w, P, 3, Bc, h, wd, u, v, r, q, partial, st);
    }
}

int tc_run_head(const float* r, const float* q, const float* scale, const DownW& w, const HeadW& hw, const balf_detector_arch& a,
                const float* blob, int Bc, int hc, int wc, float* logits, float* prob, cudaStream_t st) {
    TcPlans P;
    tc_build_plans(a, blob, &P);
    UnitGeom g{hc, wc, hc / 8, wc / 8, hc * wc / 64, Bc * (hc * wc / 64)};
    const int ntiles = (g.total_units + 1) / 2;
    const size_t smem = (size_t)TM * 256 * 4 + tc_weight_bytes(false, 0) + kTcTail;
    int grid = 0;
    if (int e = tc_launch_cfg(tc_head_kernel<256>, smem, 512, ntiles, &grid)) return e;
    {
        ProfScope ps("det_head", st);
        tc_head_kernel<256><<<grid, TM, smem, st>>>(r, q, scale, w, hw, P.head, g, a.cell, logits, prob);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

}  // namespace balf
