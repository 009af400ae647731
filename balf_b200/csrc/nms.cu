// Score-map post-processing for BALF on sm_100a: border mask, windowed NMS, greedy NMS,
// threshold, k-th-value / top-k selection, sub-pixel refinement.
//
// Reference semantics (restated in oracle/postproc.py, bit-exact on indices):
//   balf/utils/test_utils.py:34-47   remove_borders
//   balf/utils/test_utils.py:50-54   apply_nms            (windowed max compare)
//   balf/utils/test_utils.py:56-95   get_point_coordinates / find_index_higher_scores
//   balf/utils/test_utils.py:97-128  get_points_direct_from_score_map (threshold)
//   balf/utils/test_utils.py:130-168 nms_fast             (greedy, (2r+1)^2 exclusion box)
//   balf/utils/test_utils.py:170-215 soft_argmax_points   (sub-pixel)
//   demo/demo_match.py:38-57         un-pad crop, top-k by score
//
// All of this is HBM-bound byte/index work: the score map is read once (coalesced rows into a
// shared-memory halo tile), the window maximum is separable (row pass with lanes on rows, column
// pass with lanes on columns, log-step doubling networks in registers), survivors are compacted
// as 64-bit (score, ~raster) keys, and one CTA per image does radix-select + bitonic sort.
#include <algorithm>
#include <cuda.h>            // CUtensorMap (type only; the encode entry point is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

typedef unsigned long long u64;
int g_greedy_impl = 0;     // development switch (balf_debug_set key 5): 1 = the round-1 one-CTA-per-image greedy kernel

// ------------------------------------------------------------------------------------------ keys
// key = (sortable(score) << 32) | ~raster : descending key order == score desc, raster asc.
__device__ __forceinline__ uint32_t f2sortable(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float sortable2f(uint32_t s) {
    uint32_t u = s ^ ((s >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(u);
}
__device__ __forceinline__ u64 make_key(float score, uint32_t raster) {
    return ((u64)f2sortable(score) << 32) | (u64)(0xFFFFFFFFu - raster);
}
__device__ __forceinline__ uint32_t key_raster(u64 k) { return 0xFFFFFFFFu - (uint32_t)k; }
__device__ __forceinline__ float key_score(u64 k) { return sortable2f((uint32_t)(k >> 32)); }

struct MapView {
    const float* score;   // [B, Hs, Ws]
    int Hs, Ws, top, left, H, W, border;
    // border-masked score of crop pixel (y, x), both in range  (remove_borders)
    __device__ __forceinline__ float masked(int b, int y, int x) const {
        bool in = (y >= border) & (y < H - border) & (x >= border) & (x < W - border);
        return in ? __ldg(score + ((size_t)b * Hs + top + y) * Ws + left + x) : 0.0f;
    }
};

// ------------------------------------------------------------------------------------------ tile engine
// Separable window maximum over a TH x TW output tile; the window of output i covers inputs
// [i-LO, i+HI] clipped to the map (values outside the map never win: they load as the lowest T).
template <int LO, int HI, typename T>
struct Tile {
    static constexpr int TH = 32, TW = 128, SEG = 8;
    static constexpr int WIN = LO + HI + 1;
    static constexpr int IH = TH + LO + HI;
    static constexpr int NV = SEG + LO + HI;            // inputs feeding one 8-output segment
    static constexpr int VEC = 16 / (int)sizeof(T);     // elements per 128-bit shared-memory access
    static constexpr int NVV = (NV + VEC - 1) / VEC * VEC;
    static constexpr int IW_MIN = TW - SEG + NVV;       // last segment's vector reads stay in the row
    // row strides == 4 words (mod 32 banks): 128-bit accesses with lanes on rows are conflict-free
    static constexpr int IWP = (IW_MIN - VEC + 8 * VEC - 1) / (8 * VEC) * (8 * VEC) + VEC;
    static constexpr int TWP = TW + VEC;
    static constexpr int P = WIN >= 16 ? 16 : WIN >= 8 ? 8 : WIN >= 4 ? 4 : WIN >= 2 ? 2 : 1;
    static constexpr size_t smem_bytes = sizeof(T) * (size_t)IH * (IWP + TWP);
};

template <typename T> struct Vec128;
template <> struct Vec128<float> { typedef float4 type; };
template <> struct Vec128<unsigned long long> { typedef ulonglong2 type; };

template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }

// v[0..NV) -> v[i] = max(v[i..i+WIN-1]) for i < SEG   (log-step doubling, fully unrolled)
template <int LO, int HI, typename T>
__device__ __forceinline__ void window_reduce(T (&v)[Tile<LO, HI, T>::NVV]) {
    using C = Tile<LO, HI, T>;
#pragma unroll
    for (int s = 1; s < C::P; s *= 2) {
#pragma unroll
        for (int i = 0; i + s < C::NV; ++i) v[i] = tmax(v[i], v[i + s]);
    }
#pragma unroll
    for (int i = 0; i < C::SEG; ++i) v[i] = tmax(v[i], v[i + C::WIN - C::P]);
}

// Runs the tile at output origin (ty0, tx0).  load(y, x) returns the value at map position (y, x)
// (any integer position; must return the lowest T outside the map).  emit(y, x, centre, wmax) is
// called once per in-map output pixel.
template <int LO, int HI, typename T, typename Load, typename Emit>
__device__ __forceinline__ void run_tile(T* smem, int ty0, int tx0, int H, int W, Load load, Emit emit) {
    using C = Tile<LO, HI, T>;
    using V = typename Vec128<T>::type;
    T* in = smem;                                   // [IH][IWP]
    T* rm = smem + (size_t)C::IH * C::IWP;          // [IH][TWP] row maxima
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < C::IH * C::IWP; i += nt) {
        int r = i / C::IWP, c = i - r * C::IWP;
        in[i] = load(ty0 - LO + r, tx0 - LO + c);
    }
    __syncthreads();
    // row pass: a warp owns a segment column, lanes run over rows
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    constexpr int NSEG = C::TW / C::SEG;
    constexpr int RB = (C::IH + 31) / 32;
    for (int task = warp; task < NSEG * RB; task += nwarp) {
        int seg = task % NSEG, r = (task / NSEG) * 32 + lane;
        if (r < C::IH) {
            T v[C::NVV];
            const V* src = reinterpret_cast<const V*>(in + (size_t)r * C::IWP + seg * C::SEG);
#pragma unroll
            for (int j = 0; j < C::NVV / C::VEC; ++j) *reinterpret_cast<V*>(&v[j * C::VEC]) = src[j];
            window_reduce<LO, HI, T>(v);
            V* dst = reinterpret_cast<V*>(rm + (size_t)r * C::TWP + seg * C::SEG);
#pragma unroll
            for (int j = 0; j < C::SEG / C::VEC; ++j) dst[j] = *reinterpret_cast<const V*>(&v[j * C::VEC]);
        }
    }
    __syncthreads();
    // column pass: lanes run over columns, a thread owns 8 consecutive output rows
    constexpr int NRS = C::TH / C::SEG;
    for (int task = tid; task < C::TW * NRS; task += nt) {
        int c = task % C::TW, rs = task / C::TW;
        T v[C::NVV];
#pragma unroll
        for (int j = 0; j < C::NVV; ++j) v[j] = j < C::NV ? rm[(size_t)(rs * C::SEG + j) * C::TWP + c] : T(0);
        window_reduce<LO, HI, T>(v);
        int x = tx0 + c;
#pragma unroll
        for (int j = 0; j < C::SEG; ++j) {
            int y = ty0 + rs * C::SEG + j;
            if (y < H && x < W) emit(y, x, in[(size_t)(rs * C::SEG + j + LO) * C::IWP + c + LO], v[j]);
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------ windowed NMS
struct NmsWs {
    int32_t* count;   // [B] candidates per image
    int32_t* flags;   // [B] bit0 = candidate overflow
    u64* keys;        // [B][cap]
    float* alive;     // [B][H*W]  (greedy only)
    size_t cap;
};

template <int LO, int HI>
__global__ void __launch_bounds__(256) windowed_nms_kernel(MapView mv, NmsWs ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);
    using C = Tile<LO, HI, float>;
    const int b = blockIdx.z, ty0 = blockIdx.y * C::TH, tx0 = blockIdx.x * C::TW;
    auto load = [&](int y, int x) -> float {
        return (y >= 0 && y < mv.H && x >= 0 && x < mv.W) ? mv.masked(b, y, x) : kNegInf;
    };
    auto emit = [&](int y, int x, float c, float m) {
        if (c > 0.0f && c == m) {                    // survivor of apply_nms with a positive score
            unsigned slot = atomicAdd(reinterpret_cast<unsigned*>(ws.count + b), 1u);
            if (slot < ws.cap) ws.keys[(size_t)b * ws.cap + slot] = make_key(c, (uint32_t)(y * mv.W + x));
            else atomicOr(ws.flags + b, 1);
        }
    };
    run_tile<LO, HI, float>(smem, ty0, tx0, mv.H, mv.W, load, emit);
}

// ------------------------------------------------------------------------------------------ windowed NMS, 15 x 15
// The default window (nms_size 15, +-7) runs as a two-level prune instead of a dense window maximum: the dense
// separable form costs ~20 instructions per pixel, which on B200 is already the whole HBM budget (23 B/cycle/SM =
// 22 thread-instructions per fp32 pixel), while NMS survivors are ~1 % of the pixels.  One kernel, one CTA per
// 64 x 64 output tile:
//   1  the tile plus an 8-pixel halo is read once with coalesced 128-bit loads, border-masked, and kept in shared
//      memory as integers: only positive scores can survive (apply_nms keeps zeros at zero, the selection takes
//      positive keys) and for positive floats the IEEE bit pattern is monotonic, so every value is max(bits, 0);
//   2  maxima of the aligned 4 x 4 blocks (20 x 20 per tile, DPX three-input maximum);
//   3  coarse test, one thread per inner block.  With y = 4 by + r the window rows y-7 .. y+7 always contain block
//      rows by-1 .. by+1 and are contained in by-2 .. by+2, so
//        own <  max(3 x 3 blocks)  -> no pixel of the block survives            (8/9 of the blocks stop here)
//        own >= max(5 x 5 blocks)  -> every pixel equal to own survives
//        otherwise                 -> exact test against the ring blocks whose maximum exceeds own (usually 1-2);
//   4  candidates are compacted and finished with four lanes each, entirely from shared memory;
//   5  survivors are staged per CTA and appended to the image's key list with ONE global atomic (same-address
//      atomics from every warp measured ~3x slower end to end).
// The score map is read from HBM exactly once; halos (1.56x) are served by L2.
constexpr int kNmsB = 4, kNmsR = 7;
static_assert(kNmsR == 7 && kNmsB == 4, "the 3x3 / 5x5 block bounds below are derived for a +-7 window and 4x4 blocks");
constexpr int kNmsTile = 64, kNmsHalo = 8, kNmsIn = kNmsTile + 2 * kNmsHalo;       // 80 x 80 pixels in shared memory
constexpr int kNmsNB = kNmsIn / kNmsB, kNmsInner = kNmsTile / kNmsB;              // 20 x 20 blocks, 16 x 16 inner
constexpr int kNmsListCap = 512;                                                    // survivors staged per CTA

// interior test of remove_borders as two unsigned compares
struct Interior {
    int b0;
    unsigned hh, ww;      // H - 2 border, W - 2 border (0 when the border swallows the map)
    __device__ __forceinline__ bool yok(int y) const { return (unsigned)(y - b0) < hh; }
    __device__ __forceinline__ bool xok(int x) const { return (unsigned)(x - b0) < ww; }
};
__device__ __forceinline__ Interior interior_of(const MapView& mv) {
    return Interior{mv.border, (unsigned)max(mv.H - 2 * mv.border, 0), (unsigned)max(mv.W - 2 * mv.border, 0)};
}

// shared state of one tile's phases 2-5 (everything but the pixels)
struct Nms15Tile {
    int cmax[kNmsNB][kNmsNB + 1];
    uint32_t cand[kNmsInner * kNmsInner];       // lby | lbx << 8 | sure << 16
    u64 surv[kNmsListCap];
    int n_cand, n_surv, g_base;
};

// Phases 2-5 of the one-tile-per-CTA kernel on the tile at (ty0, tx0) of image b whose masked 80 x 80 window sits in px.  The
// caller has zeroed t.n_cand / t.n_surv and ended phase 1 with a CTA barrier.
__device__ __forceinline__ void nms15_phases(int (*px)[kNmsIn], Nms15Tile& t, const MapView& mv, const NmsWs& ws, int b, int ty0, int tx0) {
    const int tid = threadIdx.x, sub = tid & 3;
    // ---- 2: block maxima
    for (int i = tid; i < kNmsNB * kNmsNB; i += 256) {
        const int by = i / kNmsNB, bx = i - by * kNmsNB;
        const int4 r0 = *reinterpret_cast<const int4*>(&px[4 * by][4 * bx]);
        const int4 r1 = *reinterpret_cast<const int4*>(&px[4 * by + 1][4 * bx]);
        const int4 r2 = *reinterpret_cast<const int4*>(&px[4 * by + 2][4 * bx]);
        const int4 r3 = *reinterpret_cast<const int4*>(&px[4 * by + 3][4 * bx]);
        int m = __vimax3_s32(r0.x, r0.y, r0.z);
        m = __vimax3_s32(m, r0.w, r1.x);
        m = __vimax3_s32(m, r1.y, r1.z);
        m = __vimax3_s32(m, r1.w, r2.x);
        m = __vimax3_s32(m, r2.y, r2.z);
        m = __vimax3_s32(m, r2.w, r3.x);
        m = __vimax3_s32(m, r3.y, r3.z);
        t.cmax[by][bx] = max(m, r3.w);
    }
    __syncthreads();
    // ---- 3: coarse test, one thread per inner block
    {
        const int ly = (tid >> 4) + 2, lx = (tid & 15) + 2;
        const int own = t.cmax[ly][lx];
        if (own > 0) {
            int m3 = __vimax3_s32(t.cmax[ly - 1][lx - 1], t.cmax[ly - 1][lx], t.cmax[ly - 1][lx + 1]);
            m3 = __vimax3_s32(m3, t.cmax[ly][lx - 1], t.cmax[ly][lx + 1]);
            m3 = __vimax3_s32(m3, t.cmax[ly + 1][lx - 1], t.cmax[ly + 1][lx]);
            m3 = max(m3, t.cmax[ly + 1][lx + 1]);
            if (own >= m3) {
                int m5 = 0;
#pragma unroll
                for (int dx = -2; dx <= 2; ++dx) m5 = __vimax3_s32(m5, t.cmax[ly - 2][lx + dx], t.cmax[ly + 2][lx + dx]);
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) m5 = __vimax3_s32(m5, t.cmax[ly + dy][lx - 2], t.cmax[ly + dy][lx + 2]);
                t.cand[atomicAdd(&t.n_cand, 1)] = (uint32_t)ly | ((uint32_t)lx << 8) | (own >= m5 ? 0x10000u : 0u);
            }
        }
    }
    __syncthreads();
    // ---- 4: finish the candidates, four lanes each (sub-lane s owns row s of a block)
    const int nc = t.n_cand;
    u64* gkeys = ws.keys + (size_t)b * ws.cap;
    for (int base = 0; base < nc; base += 64) {
        const int ci = base + (tid >> 2);
        if ((ci & ~7) >= nc) continue;                      // this warp's eight candidate slots are empty (warp-uniform)
        const bool live = ci < nc;
        const uint32_t c = live ? t.cand[ci] : 0u;
        const bool sure = (c & 0x10000u) != 0;
        const int ly = live ? (int)(c & 0xFFu) : 2, lx = live ? (int)((c >> 8) & 0xFFu) : 2;
        const int own = live ? t.cmax[ly][lx] : -1;
        unsigned peaks = 0;
        {
            const int4 v = *reinterpret_cast<const int4*>(&px[4 * ly + sub][4 * lx]);
            peaks = ((v.x == own ? 1u : 0u) | (v.y == own ? 2u : 0u) | (v.z == own ? 4u : 0u) | (v.w == own ? 8u : 0u)) << (4 * sub);
        }
        peaks |= __shfl_xor_sync(0xffffffffu, peaks, 1);
        peaks |= __shfl_xor_sync(0xffffffffu, peaks, 2);
        while (__any_sync(0xffffffffu, peaks != 0)) {
            const bool act = peaks != 0;
            const int j = act ? __ffs(peaks) - 1 : 0;
            peaks &= peaks - 1;
            const int r = j >> 2, cc = j & 3;
            int wmax = 0;
            if (act && !sure) {
                // ring blocks whose maximum exceeds own; inside the window lie rows >= r+1 of block row -2, rows <= r-1
                // of block row +2, columns >= cc+1 of block column -2, columns <= cc-1 of block column +2
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int dy = k < 5 ? -2 : k < 10 ? 2 : (k - 10) / 2 - 1;
                    const int dx = k < 10 ? (k % 5) - 2 : ((k & 1) ? 2 : -2);
                    if (t.cmax[ly + dy][lx + dx] > own) {
                        const bool row_in = dy == -2 ? sub >= r + 1 : dy == 2 ? sub <= r - 1 : true;
                        if (row_in) {
                            const int4 v = *reinterpret_cast<const int4*>(&px[4 * (ly + dy) + sub][4 * (lx + dx)]);
                            const int lo = dx == -2 ? cc + 1 : 0, hi = dx == 2 ? cc - 1 : 3;
                            if (lo <= 0 && 0 <= hi) wmax = max(wmax, v.x);
                            if (lo <= 1 && 1 <= hi) wmax = max(wmax, v.y);
                            if (lo <= 2 && 2 <= hi) wmax = max(wmax, v.z);
                            if (lo <= 3 && 3 <= hi) wmax = max(wmax, v.w);
                        }
                    }
                }
            }
            wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, 1));
            wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, 2));
            if (act && sub == 0 && own >= wmax) {
                const int y = ty0 - kNmsHalo + 4 * ly + r, x = tx0 - kNmsHalo + 4 * lx + cc;
                const u64 key = make_key(__int_as_float(own), (uint32_t)(y * mv.W + x));   // own > 0: its bits are the score
                const int slot = atomicAdd(&t.n_surv, 1);
                if (slot < kNmsListCap) t.surv[slot] = key;
                else {                                       // plateau: more survivors than the staging list holds
                    const unsigned gs = atomicAdd(reinterpret_cast<unsigned*>(ws.count + b), 1u);
                    if (gs < ws.cap) gkeys[gs] = key; else atomicOr(ws.flags + b, 1);
                }
            }
        }
    }
    __syncthreads();
    // ---- 5: one append per CTA
    const int ns = min(t.n_surv, kNmsListCap);
    if (tid == 0 && ns > 0) t.g_base = (int)atomicAdd(reinterpret_cast<unsigned*>(ws.count + b), (unsigned)ns);
    __syncthreads();
    for (int i = tid; i < ns; i += 256) {
        const size_t gs = (size_t)t.g_base + i;
        if (gs < ws.cap) gkeys[gs] = t.surv[i]; else atomicOr(ws.flags + b, 1);
    }
}

// One CTA per tile, the window loaded with per-thread 128-bit loads: maps the TMA cannot describe (row pitch or base not a
// multiple of 16 bytes, maps smaller than a window) and the development switch (balf_debug_set key 7).
__global__ void __launch_bounds__(256) nms15_kernel(MapView mv, NmsWs ws, int vec_ok) {
    __shared__ __align__(128) int px[kNmsIn][kNmsIn];
    __shared__ Nms15Tile t;
    const int b = blockIdx.z, tx0 = blockIdx.x * kNmsTile, ty0 = blockIdx.y * kNmsTile;
    const int tid = threadIdx.x;
    const Interior in = interior_of(mv);
    if (tid == 0) { t.n_cand = 0; t.n_surv = 0; }
    // ---- 1: tile + halo -> shared memory
    const int* img = reinterpret_cast<const int*>(mv.score) + ((size_t)b * mv.Hs + mv.top) * mv.Ws + mv.left;   // crop origin, raw bits
    // (all of a thread's loads are issued before the first one is consumed: one HBM round trip per CTA, not seven)
    constexpr int kQuads = kNmsIn * (kNmsIn / 4), kIter = (kQuads + 255) / 256;
    int4 v[kIter];
    // Tiles whose 80 x 80 input window lies inside the border-masked interior (80 % of them at 480 x 640) skip every
    // per-pixel bounds / border test
    const bool inner = vec_ok && in.yok(ty0 - kNmsHalo) && in.yok(ty0 - kNmsHalo + kNmsIn - 1) &&
                       in.xok(tx0 - kNmsHalo) && in.xok(tx0 - kNmsHalo + kNmsIn - 1);
    if (inner) {
        const int* base = img + (size_t)(ty0 - kNmsHalo) * mv.Ws + (tx0 - kNmsHalo);
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
            const int i = tid + 256 * k;
            const int row = i / (kNmsIn / 4), c4 = i - row * (kNmsIn / 4);
            v[k] = make_int4(0, 0, 0, 0);
            if (i < kQuads) v[k] = __ldg(reinterpret_cast<const int4*>(base + (size_t)row * mv.Ws + 4 * c4));
        }
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
            const int i = tid + 256 * k;
            const int row = i / (kNmsIn / 4), c4 = i - row * (kNmsIn / 4);
            if (i < kQuads)
                *reinterpret_cast<int4*>(&px[row][4 * c4]) = make_int4(max(v[k].x, 0), max(v[k].y, 0), max(v[k].z, 0), max(v[k].w, 0));
        }
    } else {
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
        const int i = tid + 256 * k;
        const int row = i / (kNmsIn / 4), c4 = i - row * (kNmsIn / 4);
        const int gy = ty0 - kNmsHalo + row, gx = tx0 - kNmsHalo + 4 * c4;
        v[k] = make_int4(0, 0, 0, 0);
        if (i < kQuads && in.yok(gy) && gx + 3 >= in.b0 && gx < mv.W - in.b0) {
            const int* src = img + (size_t)gy * mv.Ws + gx;
            if (vec_ok && gx >= 0 && gx + 3 < mv.W) v[k] = __ldg(reinterpret_cast<const int4*>(src));
            else {
                if (gx >= 0 && gx < mv.W) v[k].x = __ldg(src);
                if (gx + 1 >= 0 && gx + 1 < mv.W) v[k].y = __ldg(src + 1);
                if (gx + 2 >= 0 && gx + 2 < mv.W) v[k].z = __ldg(src + 2);
                if (gx + 3 >= 0 && gx + 3 < mv.W) v[k].w = __ldg(src + 3);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
        const int i = tid + 256 * k;
        const int row = i / (kNmsIn / 4), c4 = i - row * (kNmsIn / 4);
        const int gx = tx0 - kNmsHalo + 4 * c4;
        int4 o;
        o.x = in.xok(gx) ? max(v[k].x, 0) : 0;
        o.y = in.xok(gx + 1) ? max(v[k].y, 0) : 0;
        o.z = in.xok(gx + 2) ? max(v[k].z, 0) : 0;
        o.w = in.xok(gx + 3) ? max(v[k].w, 0) : 0;
        if (i < kQuads) *reinterpret_cast<int4*>(&px[row][4 * c4]) = o;
    }
    }
    __syncthreads();
    nms15_phases(px, t, mv, ws, b, ty0, tx0);
}

// The TMA form (every map whose rows and crop the tensor map can describe): PERSISTENT CTAs, two window buffers each.
//  * The 80 x 80 window of a tile arrives by one cp.async.bulk.tensor of the box of a 3-D tensor map over the score maps
//    [B, Hs, Ws] into one of two shared-memory buffers, completion on that buffer's mbarrier -- no thread instruction, no index
//    arithmetic; the window of the CTA's next tile is in flight while the current one is worked on.  Windows that reach over the
//    border-masked interior, the crop or the map come in as raw values (zeros outside the map: the TMA's out-of-bounds fill) and
//    are masked in place while the block maxima are taken.  The raw fp32 bits are used as they are: scores are >= 0 (softmax
//    probabilities), and a negative value orders below zero as an integer, i.e. it can never survive -- the same outcome as the
//    max(bits, 0) of the manual path.
//  * ONE CTA barrier per tile.  The one-tile-per-CTA form is a chain of five barrier-separated phases, most of them a few dozen
//    instructions deep, and spent 40 % of its stall samples at those barriers (ncu, round 2).  Here the CTA takes the block maxima
//    together (phase 2, one barrier), then every warp works alone on the two inner block rows 2w+2, 2w+3 it owns: coarse test
//    with one lane per block, its candidates finished four lanes each, __syncwarp only -- and moves on to the next tile while other
//    warps are still busy.  Survivors are staged per tile; one warp (a different one every tile) collects the other
//    seven on a named barrier (bar.arrive / bar.sync), appends them to the image's key list with one global atomic and requests
//    the window of the tile after next into the buffer just freed.
//    (Measured alternatives: warps that also take their own block maxima of the six block rows they read -- no barrier at all, 2.4x
//    the phase-2 work -- 45 us against 37 (scripts/nms_bench.py); sixteen lanes per candidate with warp-group reductions instead of the serial walk over
//    the ring blocks: 50 us; four ring blocks per sub-lane (a four-step instead of a sixteen-step walk): 42 us -- every variant
//    that shortens a dependency chain at the price of more instructions lost.  Phase-skipping runs put the kernel's time at:
//    block maxima + barrier 9 us, coarse test 3, candidates 10, append 2, loop / hand-over 7; the window loads alone take
//    19 us and hide behind the rest.)
constexpr int kNmsPxBytes = kNmsIn * kNmsIn * 4;
constexpr int kNmsSurvCap = 64;               // survivors staged per tile (a 64 x 64 tile holds at most 64 without plateaus)
struct Nms15W {
    int cmax[2][kNmsNB][kNmsNB + 1];          // per window buffer: a warp may be one tile ahead of the slowest one
    u64 surv[2][kNmsSurvCap];
    unsigned char cl[8][32];                  // per warp: lane (| 0x80: sure) of each candidate block
    int n_surv[2];
    int info[2][4];                           // tile in a buffer: image, ty0, tx0, window inside the interior
    unsigned long long full[2];
};
__device__ __forceinline__ uint32_t nms_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 5) nms15_tma_kernel(MapView mv, NmsWs ws, const __grid_constant__ CUtensorMap tmap,
                                                        int tiles_x, int tiles_y, int ntiles) {
    extern __shared__ __align__(128) unsigned char nms_dyn[];
    __shared__ __align__(16) Nms15W S;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, sub = lane & 3;
    const Interior in = interior_of(mv);
    const int per_img = tiles_x * tiles_y, stride = (int)gridDim.x;
    auto request = [&](int tile, int buf) {                 // one thread: the window of `tile` -> buffer `buf`
        const int b = tile / per_img, r = tile - b * per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
        const int ty0 = ty * kNmsTile, tx0 = tx * kNmsTile;
        S.info[buf][0] = b; S.info[buf][1] = ty0; S.info[buf][2] = tx0;
        S.info[buf][3] = in.yok(ty0 - kNmsHalo) && in.yok(ty0 - kNmsHalo + kNmsIn - 1) &&
                         in.xok(tx0 - kNmsHalo) && in.xok(tx0 - kNmsHalo + kNmsIn - 1);
        const uint32_t bar = nms_smem_u32(&S.full[buf]);
        // (the arrive releases the info words to the warps that acquire the phase in their try_wait)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)kNmsPxBytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(nms_smem_u32(nms_dyn + buf * kNmsPxBytes)), "l"(&tmap), "r"(mv.left + tx0 - kNmsHalo),
                        "r"(mv.top + ty0 - kNmsHalo), "r"(b), "r"(bar)
                     : "memory");
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(nms_smem_u32(&S.full[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(nms_smem_u32(&S.full[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.n_surv[0] = S.n_surv[1] = 0;
        if ((int)blockIdx.x < ntiles) request((int)blockIdx.x, 0);
        if ((int)blockIdx.x + stride < ntiles) request((int)blockIdx.x + stride, 1);
    }
    __syncthreads();                                        // mbarriers and counters are initialised
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += stride, ++it) {
        const int buf = it & 1;
        int (*px)[kNmsIn] = reinterpret_cast<int (*)[kNmsIn]>(nms_dyn + buf * kNmsPxBytes);
        int (*cm)[kNmsNB + 1] = S.cmax[buf];
        {
            const uint32_t bar = nms_smem_u32(&S.full[buf]), parity = (uint32_t)(it >> 1) & 1u;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            }
        }
        const int b = S.info[buf][0], ty0 = S.info[buf][1], tx0 = S.info[buf][2];
        const bool inner = S.info[buf][3] != 0;
        // ---- 2: maxima of the 20 x 20 blocks, the whole CTA (masking the window in place where it leaves the interior)
        for (int i = tid; i < kNmsNB * kNmsNB; i += 256) {
            const int by = i / kNmsNB, bx = i - by * kNmsNB;
            int4 r0 = *reinterpret_cast<const int4*>(&px[4 * by][4 * bx]);
            int4 r1 = *reinterpret_cast<const int4*>(&px[4 * by + 1][4 * bx]);
            int4 r2 = *reinterpret_cast<const int4*>(&px[4 * by + 2][4 * bx]);
            int4 r3 = *reinterpret_cast<const int4*>(&px[4 * by + 3][4 * bx]);
            if (!inner) {
                const int gy = ty0 - kNmsHalo + 4 * by, gx = tx0 - kNmsHalo + 4 * bx;
                const bool y0 = in.yok(gy), y1 = in.yok(gy + 1), y2 = in.yok(gy + 2), y3 = in.yok(gy + 3);
                const bool x0 = in.xok(gx), x1 = in.xok(gx + 1), x2 = in.xok(gx + 2), x3 = in.xok(gx + 3);
                if (!(y0 & y1 & y2 & y3 & x0 & x1 & x2 & x3)) {        // a block on (or beyond) the interior's boundary
                    r0 = make_int4(y0 & x0 ? r0.x : 0, y0 & x1 ? r0.y : 0, y0 & x2 ? r0.z : 0, y0 & x3 ? r0.w : 0);
                    r1 = make_int4(y1 & x0 ? r1.x : 0, y1 & x1 ? r1.y : 0, y1 & x2 ? r1.z : 0, y1 & x3 ? r1.w : 0);
                    r2 = make_int4(y2 & x0 ? r2.x : 0, y2 & x1 ? r2.y : 0, y2 & x2 ? r2.z : 0, y2 & x3 ? r2.w : 0);
                    r3 = make_int4(y3 & x0 ? r3.x : 0, y3 & x1 ? r3.y : 0, y3 & x2 ? r3.z : 0, y3 & x3 ? r3.w : 0);
                    *reinterpret_cast<int4*>(&px[4 * by][4 * bx]) = r0;
                    *reinterpret_cast<int4*>(&px[4 * by + 1][4 * bx]) = r1;
                    *reinterpret_cast<int4*>(&px[4 * by + 2][4 * bx]) = r2;
                    *reinterpret_cast<int4*>(&px[4 * by + 3][4 * bx]) = r3;
                }
            }
            int m = __vimax3_s32(r0.x, r0.y, r0.z);
            m = __vimax3_s32(m, r0.w, r1.x);
            m = __vimax3_s32(m, r1.y, r1.z);
            m = __vimax3_s32(m, r1.w, r2.x);
            m = __vimax3_s32(m, r2.y, r2.z);
            m = __vimax3_s32(m, r2.w, r3.x);
            m = __vimax3_s32(m, r3.y, r3.z);
            cm[by][bx] = max(m, r3.w);
        }
        __syncthreads();                                    // the tile's only CTA barrier
        // ---- 3: coarse test, one lane per inner block; warp w owns the inner block rows 2w + 2, 2w + 3
        unsigned cmask;
        {
            const int rr = 2 * w + 2 + (lane >> 4), lx = 2 + (lane & 15);
            const int own = cm[rr][lx];
            bool is_c = false, sure = false;
            if (own > 0) {
                int m3 = __vimax3_s32(cm[rr - 1][lx - 1], cm[rr - 1][lx], cm[rr - 1][lx + 1]);
                m3 = __vimax3_s32(m3, cm[rr][lx - 1], cm[rr][lx + 1]);
                m3 = __vimax3_s32(m3, cm[rr + 1][lx - 1], cm[rr + 1][lx]);
                m3 = max(m3, cm[rr + 1][lx + 1]);
                if (own >= m3) {
                    int m5 = 0;
#pragma unroll
                    for (int dx = -2; dx <= 2; ++dx) m5 = __vimax3_s32(m5, cm[rr - 2][lx + dx], cm[rr + 2][lx + dx]);
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) m5 = __vimax3_s32(m5, cm[rr + dy][lx - 2], cm[rr + dy][lx + 2]);
                    is_c = true;
                    sure = own >= m5;
                }
            }
            cmask = __ballot_sync(0xffffffffu, is_c);
            if (is_c) S.cl[w][__popc(cmask & ((1u << lane) - 1u))] = (unsigned char)(lane | (sure ? 0x80 : 0));
        }
        __syncwarp();
        // ---- 4: finish the candidates, four lanes each (sub-lane s owns row s of a block)
        const int nc = __popc(cmask);
        u64* gkeys = ws.keys + (size_t)b * ws.cap;
        for (int base = 0; base < nc; base += 8) {
            const int ci = base + (lane >> 2);
            const bool live = ci < nc;
            const unsigned c = live ? S.cl[w][ci] : 0u;
            const bool sure = (c & 0x80u) != 0;
            const int rr = 2 * w + 2 + (int)((c >> 4) & 1u), lx = 2 + (int)(c & 15u), ly = rr;     // block row / column of the tile
            const int own = live ? cm[rr][lx] : -1;
            unsigned peaks = 0;
            {
                const int4 v = *reinterpret_cast<const int4*>(&px[4 * ly + sub][4 * lx]);
                peaks = ((v.x == own ? 1u : 0u) | (v.y == own ? 2u : 0u) | (v.z == own ? 4u : 0u) | (v.w == own ? 8u : 0u)) << (4 * sub);
            }
            peaks |= __shfl_xor_sync(0xffffffffu, peaks, 1);
            peaks |= __shfl_xor_sync(0xffffffffu, peaks, 2);
            while (__any_sync(0xffffffffu, peaks != 0)) {
                const bool act = peaks != 0;
                const int j = act ? __ffs(peaks) - 1 : 0;
                peaks &= peaks - 1;
                const int r = j >> 2, cc = j & 3;
                int wmax = 0;
                if (act && !sure) {
                    // ring blocks whose maximum exceeds own; inside the window lie rows >= r+1 of block row -2, rows <= r-1
                    // of block row +2, columns >= cc+1 of block column -2, columns <= cc-1 of block column +2
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int dy = k < 5 ? -2 : k < 10 ? 2 : (k - 10) / 2 - 1;
                        const int dx = k < 10 ? (k % 5) - 2 : ((k & 1) ? 2 : -2);
                        if (cm[rr + dy][lx + dx] > own) {
                            const bool row_in = dy == -2 ? sub >= r + 1 : dy == 2 ? sub <= r - 1 : true;
                            if (row_in) {
                                const int4 v = *reinterpret_cast<const int4*>(&px[4 * (ly + dy) + sub][4 * (lx + dx)]);
                                const int lo = dx == -2 ? cc + 1 : 0, hi = dx == 2 ? cc - 1 : 3;
                                if (lo <= 0 && 0 <= hi) wmax = max(wmax, v.x);
                                if (lo <= 1 && 1 <= hi) wmax = max(wmax, v.y);
                                if (lo <= 2 && 2 <= hi) wmax = max(wmax, v.z);
                                if (lo <= 3 && 3 <= hi) wmax = max(wmax, v.w);
                            }
                        }
                    }
                }
                wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, 1));
                wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, 2));
                if (act && sub == 0 && own >= wmax) {
                    const int y = ty0 - kNmsHalo + 4 * ly + r, x = tx0 - kNmsHalo + 4 * lx + cc;
                    const u64 key = make_key(__int_as_float(own), (uint32_t)(y * mv.W + x));   // own > 0: its bits are the score
                    const int slot = atomicAdd(&S.n_surv[buf], 1);
                    if (slot < kNmsSurvCap) S.surv[buf][slot] = key;
                    else {                                   // plateau: more survivors than the staging list holds
                        const unsigned gs = atomicAdd(reinterpret_cast<unsigned*>(ws.count + b), 1u);
                        if (gs < ws.cap) gkeys[gs] = key; else atomicOr(ws.flags + b, 1);
                    }
                }
            }
        }
        __syncwarp();
        // ---- 5: every warp signs the tile off on the buffer's named barrier without waiting (bar.arrive); one warp -- a different
        // one every tile -- waits there for all eight (bar.sync), appends the tile's survivors with one global atomic and
        // refills the buffer.  The others are already in the next tile.  (A shared-memory counter with fences -- "the last warp
        // to arrive appends" -- measured 1 us faster per launch, but compute-sanitizer's racecheck cannot follow it.)
        if (w != (it & 7)) {
            if (buf) asm volatile("bar.arrive 2, 256;" ::: "memory"); else asm volatile("bar.arrive 1, 256;" ::: "memory");
        } else {
            if (buf) asm volatile("bar.sync 2, 256;" ::: "memory"); else asm volatile("bar.sync 1, 256;" ::: "memory");
            const int ns = min(*reinterpret_cast<volatile int*>(&S.n_surv[buf]), kNmsSurvCap);
            unsigned gb = 0;
            if (lane == 0 && ns > 0) gb = atomicAdd(reinterpret_cast<unsigned*>(ws.count + b), (unsigned)ns);
            gb = __shfl_sync(0xffffffffu, gb, 0);
            for (int i = lane; i < ns; i += 32) {
                const size_t gs = (size_t)gb + i;
                const u64 key = *reinterpret_cast<volatile u64*>(&S.surv[buf][i]);
                if (gs < ws.cap) gkeys[gs] = key; else atomicOr(ws.flags + b, 1);
            }
            __syncwarp();
            if (lane == 0) {
                S.n_surv[buf] = 0;
                // the generic-proxy accesses of the window (every warp has left it) are ordered before the TMA's write
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (tile + 2 * stride < ntiles) request(tile + 2 * stride, buf);
            }
        }
    }
}

// any window size: one thread per pixel, window read through L1/L2 (small maps, unusual sizes)
__global__ void windowed_nms_generic_kernel(MapView mv, NmsWs ws, int lo, int hi) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= mv.W) return;
    float c = mv.masked(b, y, x);
    if (!(c > 0.0f)) return;
    int y0 = max(y - lo, 0), y1 = min(y + hi, mv.H - 1), x0 = max(x - lo, 0), x1 = min(x + hi, mv.W - 1);
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx)
            if (mv.masked(b, yy, xx) > c) return;
    unsigned slot = atomicAdd(reinterpret_cast<unsigned*>(ws.count + b), 1u);
    if (slot < ws.cap) ws.keys[(size_t)b * ws.cap + slot] = make_key(c, (uint32_t)(y * mv.W + x));
    else atomicOr(ws.flags + b, 1);
}

// exclusion footprint of a kept pixel: rows |dy| <= reach, half width hw[|dy|] (a square of radius r for nms_fast; the
// offsets whose box IoU exceeds the threshold for box_nms)
constexpr int kFpMax = 64;
struct Footprint {
    int S;              // cell size of the cell kernels (<= 16)
    int reach;          // largest |dy| of the footprint
    int hw[kFpMax + 1]; // half width at |dy| (-1: empty row)
};
__device__ __forceinline__ bool fp_inside(const Footprint& f, int dy, int dx) {
    const int ay = abs(dy);
    return ay <= f.reach && abs(dx) <= f.hw[ay];
}

// ------------------------------------------------------------------------------------------ greedy NMS
// One CTA per image.  Rounds: every alive pixel whose 64-bit key is the maximum of the alive keys in
// its (2r+1)^2 window is kept; everything alive within r of a kept pixel dies; repeat until nothing
// is alive.  Equal to the sequential walk of nms_fast.
__global__ void greedy_init_kernel(MapView mv, NmsWs ws, float thr) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= mv.W) return;
    float s = mv.masked(b, y, x);
    ws.alive[(size_t)b * mv.H * mv.W + (size_t)y * mv.W + x] = (s >= thr) ? s : kNegInf;
}

template <int R>
__device__ __forceinline__ void greedy_round_tiles(u64* smem, float* alive, int H, int W, u64* kept, int cap,
                                                   int* n_kept, int* n_new, int* sh_new, int sh_cap) {
    using C = Tile<R, R, u64>;
    const int tiles_x = (W + C::TW - 1) / C::TW, tiles_y = (H + C::TH - 1) / C::TH;
    for (int t = 0; t < tiles_x * tiles_y; ++t) {
        int ty0 = (t / tiles_x) * C::TH, tx0 = (t % tiles_x) * C::TW;
        auto load = [&](int y, int x) -> u64 {
            if (y < 0 || y >= H || x < 0 || x >= W) return 0ull;
            float a = alive[(size_t)y * W + x];
            return a > kNegInf ? make_key(a, (uint32_t)(y * W + x)) : 0ull;
        };
        auto emit = [&](int y, int x, u64 c, u64 m) {
            if (c != 0ull && c == m) {
                int slot = atomicAdd(n_kept, 1);
                if (slot < cap) kept[slot] = c;
                int s2 = atomicAdd(n_new, 1);
                if (s2 < sh_cap) sh_new[s2] = y * W + x;
            }
        };
        run_tile<R, R, u64>(smem, ty0, tx0, H, W, load, emit);
    }
}

__device__ __forceinline__ void greedy_round_generic(float* alive, int H, int W, const Footprint& f, u64* kept, int cap,
                                                     int* n_kept, int* n_new, int* sh_new, int sh_cap) {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        float a = alive[i];
        if (!(a > kNegInf)) continue;
        int y = i / W, x = i - y * W;
        u64 me = make_key(a, (uint32_t)i);
        int y0 = max(y - f.reach, 0), y1 = min(y + f.reach, H - 1);
        bool top = true;
        for (int yy = y0; yy <= y1 && top; ++yy) {
            const int hw = f.hw[abs(yy - y)];
            for (int xx = max(x - hw, 0); xx <= min(x + hw, W - 1); ++xx) {
                float q = alive[(size_t)yy * W + xx];
                if (q > kNegInf && make_key(q, (uint32_t)(yy * W + xx)) > me) { top = false; break; }
            }
        }
        if (top) {
            int slot = atomicAdd(n_kept, 1);
            if (slot < cap) kept[slot] = me;
            int s2 = atomicAdd(n_new, 1);
            if (s2 < sh_cap) sh_new[s2] = i;
        }
    }
}
// one warp clears the footprint of a kept pixel
__device__ __forceinline__ void greedy_suppress(float* alive, int H, int W, const Footprint& f, int p) {
    const int y = p / W, x = p - y * W;
    for (int yy = max(y - f.reach, 0); yy <= min(y + f.reach, H - 1); ++yy) {
        const int hw = f.hw[abs(yy - y)];
        for (int xx = max(x - hw, 0) + (threadIdx.x & 31); xx <= min(x + hw, W - 1); xx += 32) alive[(size_t)yy * W + xx] = kNegInf;
    }
}

constexpr int kGreedyNewCap = 4096;   // kept pixels per round staged in shared memory

template <int R>   // R > 0: tile engine with square radius R;  R == 0: any footprint, direct window scans
__global__ void __launch_bounds__(256) greedy_nms_kernel(NmsWs ws, int H, int W, Footprint fp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int n_kept, n_new;
    __shared__ int sh_new[kGreedyNewCap];
    __shared__ Footprint f;
    const int b = blockIdx.x;
    float* alive = ws.alive + (size_t)b * H * W;
    u64* kept = ws.keys + (size_t)b * ws.cap;
    if (threadIdx.x == 0) { n_kept = 0; f = fp; }
    __syncthreads();
    for (int round = 0; round < (1 << 24); ++round) {      // bounded: every round keeps >= 1 pixel or ends
        if (threadIdx.x == 0) n_new = 0;
        __syncthreads();
        if constexpr (R > 0) greedy_round_tiles<(R > 0 ? R : 1)>(reinterpret_cast<u64*>(smem_raw), alive, H, W, kept, (int)ws.cap,
                                                       &n_kept, &n_new, sh_new, kGreedyNewCap);
        else greedy_round_generic(alive, H, W, f, kept, (int)ws.cap, &n_kept, &n_new, sh_new, kGreedyNewCap);
        __syncthreads();
        const int nn = n_new;
        if (nn == 0) break;
        if (nn <= kGreedyNewCap) {
            for (int i = threadIdx.x >> 5; i < nn; i += blockDim.x >> 5) greedy_suppress(alive, H, W, f, sh_new[i]);
        } else {
            // more new keeps than the staging list holds (tiny footprint): re-derive them from the
            // kept list itself -- the last nn entries were appended this round.
            const int total = min(n_kept, (int)ws.cap);
            for (int i = total - nn + (threadIdx.x >> 5); i < total; i += blockDim.x >> 5) {
                if (i < 0) continue;
                greedy_suppress(alive, H, W, f, (int)key_raster(kept[i]));
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ws.count[b] = min(n_kept, (int)ws.cap);
        if (n_kept > (int)ws.cap) ws.flags[b] |= 1;
    }
}

// ------------------------------------------------------------------------------------------ greedy NMS on cells
// The same greedy result (nms_fast, test_utils.py:130-168; torchvision.ops.nms for box_nms, repeatability_tools.py:227-255)
// with work proportional to the ALIVE pixels instead of whole-image window maxima, spread over the whole GPU.
// The map is cut into S x S cells, S chosen so that (1) any two pixels of one cell exclude each other and (2) the exclusion
// footprint of a pixel stays inside its 3 x 3 cell neighbourhood (square radius r: S = r + 1).  Then
//   * at most one pixel per cell is ever kept, and a pixel can only be kept while it is the maximum alive key of its cell;
//   * per round, one warp per cell: the cell maximum p is KEPT when no alive pixel with a larger key lies inside its footprint:
//     the eight neighbour maxima decide almost every case (smaller -> cannot interfere; larger and inside the footprint ->
//     blocked), the rest scans the alive pixels of that neighbour; a kept pixel clears the alive bits of its footprint
//     (atomicAnd on 256-bit cell masks) and marks the touched cells dirty;
//   * dirty cells recompute their maximum from their alive pixels only (first round: threshold + border mask).
// Every decision is monotone -- alive bits only ever clear, a stale cell maximum is an upper bound -- so concurrent warps
// can only be conservative (a pixel waits one more round), never wrong; two cell maxima inside each other's footprint
// always see each other, so exactly the larger one proceeds.  Rounds run as kernel pairs over all cells of the batch; a
// one-CTA-per-image kernel finishes pathological maps (long monotone ramps need one round per kept pixel).
struct CellWs {
    uint32_t* alive;    // [B][cells][8]   bit ly * 16 + lx of cell pixel (ly, lx)
    u64* cmax;          // [B][cells]      maximum alive key, 0 = none
    u64* kept;          // [B][cells]      the cell's kept key (at most one per cell), 0 = none; compacted into the key list at the end
    int* dirty;         // [B][cells]
    int cx, cy;         // cells per row / column
    int vec_ok;         // S = 16 and 16-byte aligned rows: interior cells of the first pass use 128-bit loads
};
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
// this lane's 8 pixels of a cell: row ly = lane >> 1, columns (lane & 1) * 8 .. + 7
// (re)compute the maximum alive key of cell c (one warp); init: alive = masked score >= thr
__device__ __forceinline__ void cell_refresh(const MapView& mv, const CellWs& cw, const Footprint& f, int b, int c, bool init,
                                             float thr) {
    const int lane = threadIdx.x & 31;
    const size_t cell = (size_t)b * cw.cx * cw.cy + c;
    if (!init && lane == 0) cw.dirty[cell] = 0;
    const int cyi = c / cw.cx, cxi = c - cyi * cw.cx;
    const int ly = lane >> 1, lx0 = (lane & 1) * 8;
    const int y = cyi * f.S + ly;
    uint32_t word = 0;
    if (!init && lane < 8) word = cw.alive[cell * 8 + lane];
    uint32_t bits = init ? 0xFFu : (__shfl_sync(0xffffffffu, word, ly >> 1) >> ((ly & 1) * 16 + lx0)) & 0xFFu;
    u64 best = 0;
    uint32_t mine = 0;
    // first pass, cells that lie inside the border-masked interior: two 128-bit loads per lane, no per-pixel tests
    const bool interior = init && cw.vec_ok && cyi * 16 >= mv.border && cyi * 16 + 16 <= mv.H - mv.border &&
                          cxi * 16 >= mv.border && cxi * 16 + 16 <= mv.W - mv.border;
    if (interior) {
        const float4* src = reinterpret_cast<const float4*>(mv.score + ((size_t)b * mv.Hs + mv.top + y) * mv.Ws + mv.left + cxi * 16 + lx0);
        const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
        const float sc8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        const uint32_t r0 = (uint32_t)(y * mv.W + cxi * 16 + lx0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (sc8[j] >= thr) {
                mine |= 1u << j;
                const u64 k = make_key(sc8[j], r0 + j);
                best = k > best ? k : best;
            }
        }
    } else if (ly < f.S && y < mv.H && bits) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int lx = lx0 + j, x = cxi * f.S + lx;
            if (((bits >> j) & 1u) && lx < f.S && x < mv.W) {
                const float sc = mv.masked(b, y, x);
                if (!init || sc >= thr) {
                    mine |= 1u << j;
                    const u64 k = make_key(sc, (uint32_t)(y * mv.W + x));
                    best = k > best ? k : best;
                }
            }
        }
    }
    best = warp_max_u64(best);
    if (init) {
        uint32_t w = mine << ((ly & 1) * 16 + lx0);
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((lane & 3) == 0) cw.alive[cell * 8 + (lane >> 2)] = w;
    }
    if (lane == 0) cw.cmax[cell] = best;
}
// one THREAD: can the maximum of cell c be kept?  0 = no (empty cell, or a larger neighbour maximum lies inside its footprint);
// otherwise 1 | (mask of the neighbours whose larger maximum lies outside the footprint: their alive pixels must be scanned) << 1
__device__ __forceinline__ uint32_t cell_classify(const CellWs& cw, const Footprint& f, int W, int b, int c) {
    const size_t cell0 = (size_t)b * cw.cx * cw.cy;
    const u64 p = cw.cmax[cell0 + c];
    if (p == 0ull) return 0u;
    const int cyi = c / cw.cx, cxi = c - cyi * cw.cx;
    const int pr = (int)key_raster(p), py = pr / W, px = pr - py * W;
    u64 m[9];
#pragma unroll
    for (int n = 0; n < 9; ++n) {
        const int ny = cyi + n / 3 - 1, nx = cxi + n % 3 - 1;
        m[n] = (n != 4 && ny >= 0 && ny < cw.cy && nx >= 0 && nx < cw.cx) ? cw.cmax[cell0 + ny * cw.cx + nx] : 0ull;
    }
    uint32_t slow = 0;
#pragma unroll
    for (int n = 0; n < 9; ++n) {
        if (m[n] > p) {
            const int qr = (int)key_raster(m[n]), qy = qr / W, qx = qr - qy * W;
            if (fp_inside(f, qy - py, qx - px)) return 0u;
            slow |= 1u << n;
        }
    }
    return 1u | (slow << 1);
}
// one WARP: scan the flagged neighbours for an alive pixel with a larger key inside the footprint; none -> keep the cell maximum
// and clear its footprint
__device__ __forceinline__ void cell_keep(const MapView& mv, const CellWs& cw, const Footprint& f, const NmsWs& ws, int b, int c,
                                          uint32_t slow) {
    const int lane = threadIdx.x & 31;
    const size_t cell0 = (size_t)b * cw.cx * cw.cy;
    const u64 p = cw.cmax[cell0 + c];
    const int cyi = c / cw.cx, cxi = c - cyi * cw.cx;
    const int pr = (int)key_raster(p), py = pr / mv.W, px = pr - py * mv.W;
    while (slow) {
        const int n = __ffs(slow) - 1;
        slow &= slow - 1;
        const int ny = cyi + n / 3 - 1, nx = cxi + n % 3 - 1;
        const size_t nc = cell0 + (size_t)ny * cw.cx + nx;
        uint32_t word = lane < 8 ? cw.alive[nc * 8 + lane] : 0u;
        const int ly = lane >> 1, lx0 = (lane & 1) * 8, y = ny * f.S + ly;
        const uint32_t bits = (__shfl_sync(0xffffffffu, word, ly >> 1) >> ((ly & 1) * 16 + lx0)) & 0xFFu;
        bool hit = false;
        if (bits && abs(y - py) <= f.reach) {
            const int hwid = f.hw[abs(y - py)];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = nx * f.S + lx0 + j;
                if (((bits >> j) & 1u) && abs(x - px) <= hwid) {
                    const u64 k = make_key(mv.masked(b, y, x), (uint32_t)(y * mv.W + x));
                    hit = hit || k > p;
                }
            }
        }
        if (__any_sync(0xffffffffu, hit)) return;
    }
    if (lane == 0) cw.kept[cell0 + c] = p;             // no atomics: the slot is the cell (same-address atomics on the image's
                                                       // counter serialised the ~350 keeps of a first round)
    for (int t = lane; t < 72; t += 32) {
        const int n = t >> 3, j = t & 7;
        const int ny = cyi + n / 3 - 1, nx = cxi + n % 3 - 1;
        if (ny < 0 || ny >= cw.cy || nx < 0 || nx >= cw.cx) continue;
        uint32_t mask = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ly = 2 * j + h, y = ny * f.S + ly, ady = abs(y - py);
            if (ly >= f.S || ady > f.reach) continue;
            const int hwid = f.hw[ady];
            if (hwid < 0) continue;
            const int lo = max(px - hwid - nx * f.S, 0), hi = min(px + hwid - nx * f.S, f.S - 1);
            if (lo <= hi) mask |= (((1u << (hi - lo + 1)) - 1u) << lo) << (h * 16);
        }
        if (mask) {
            const size_t nc = cell0 + (size_t)ny * cw.cx + nx;
            atomicAnd(&cw.alive[nc * 8 + j], ~mask);
            cw.dirty[nc] = 1;
        }
    }
}
// a warp takes kCpw consecutive cells [base, base + kCpw) of the flattened (image, cell) range: the cheap per-cell tests run
// one thread per cell (independent load chains in flight), the heavy steps (about one cell in five in the first round) run one
// warp per cell; kCpw = 8 spreads those over every warp of the grid (32 measured 1.5x slower: ten serial keeps per warp)
constexpr int kCpw = 8;
__device__ __forceinline__ void cells_refresh32(const MapView& mv, const CellWs& cw, const Footprint& f, int base, int total, int nc) {
    const int lane = threadIdx.x & 31, i = base + lane;
    unsigned todo = __ballot_sync(0xffffffffu, lane < kCpw && i < total && cw.dirty[i] != 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int ci = base + src, b = ci / nc;
        cell_refresh(mv, cw, f, b, ci - b * nc, false, 0.f);
    }
}
__device__ __forceinline__ void cells_decide32(const MapView& mv, const CellWs& cw, const Footprint& f, const NmsWs& ws, int base,
                                               int total, int nc) {
    const int lane = threadIdx.x & 31, i = base + lane;
    uint32_t cls = 0;
    if (lane < kCpw && i < total) { const int b = i / nc; cls = cell_classify(cw, f, mv.W, b, i - b * nc); }
    unsigned todo = __ballot_sync(0xffffffffu, cls != 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t slow = __shfl_sync(0xffffffffu, cls, src) >> 1;
        const int ci = base + src, b = ci / nc;
        cell_keep(mv, cw, f, ws, b, ci - b * nc, slow);
    }
}

// persistent grids: a few CTAs per SM walk all cells of the batch (launching one CTA per eight cells cost ~25 us per pass in
// CTA scheduling alone, ten times the work of a late round)
__global__ void __launch_bounds__(256) greedy_cells_init_kernel(MapView mv, CellWs cw, Footprint f, float thr, int B) {
    __shared__ Footprint fs;
    if (threadIdx.x == 0) fs = f;
    __syncthreads();
    const int nc = cw.cx * cw.cy, total = nc * B;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < total; i += gridDim.x * 8) {
        const int b = i / nc;
        cell_refresh(mv, cw, fs, b, i - b * nc, true, thr);
    }
}
__global__ void __launch_bounds__(256) greedy_cells_refresh_kernel(MapView mv, CellWs cw, Footprint f, int B) {
    __shared__ Footprint fs;
    if (threadIdx.x == 0) fs = f;
    __syncthreads();
    const int nc = cw.cx * cw.cy, total = nc * B;
    for (int base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kCpw; base < total; base += gridDim.x * 8 * kCpw)
        cells_refresh32(mv, cw, fs, base, total, nc);
}
__global__ void __launch_bounds__(256) greedy_cells_decide_kernel(MapView mv, CellWs cw, Footprint f, NmsWs ws, int B) {
    __shared__ Footprint fs;
    if (threadIdx.x == 0) fs = f;
    __syncthreads();
    const int nc = cw.cx * cw.cy, total = nc * B;
    for (int base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kCpw; base < total; base += gridDim.x * 8 * kCpw)
        cells_decide32(mv, cw, fs, ws, base, total, nc);
}
// one CTA per image: runs rounds until no cell has an alive pixel (returns at once when the multi-CTA rounds finished the job)
__global__ void __launch_bounds__(1024) greedy_cells_finish_kernel(MapView mv, CellWs cw, Footprint f, NmsWs ws) {
    __shared__ Footprint fs;
    __shared__ int alive_cells;
    if (threadIdx.x == 0) fs = f;
    const int b = blockIdx.x, nc = cw.cx * cw.cy, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int first = b * nc, last = first + nc;              // this image's slice of the flattened cell range
    for (int round = 0; round < (1 << 24); ++round) {          // every round keeps at least the largest alive key
        __syncthreads();
        if (threadIdx.x == 0) alive_cells = 0;
        __syncthreads();
        for (int base = first + warp * kCpw; base < last; base += nw * kCpw) cells_refresh32(mv, cw, fs, base, last, nc);
        __syncthreads();
        int loc = 0;
        for (int c = first + threadIdx.x; c < last; c += blockDim.x) loc += cw.cmax[c] != 0ull ? 1 : 0;
        if (loc) atomicAdd(&alive_cells, loc);
        __syncthreads();
        if (alive_cells == 0) break;
        for (int base = first + warp * kCpw; base < last; base += nw * kCpw) cells_decide32(mv, cw, fs, ws, base, last, nc);
        __threadfence_block();
    }
    // kept keys of the cells -> the image's key list (order is irrelevant: the select / sort kernel orders by key)
    __shared__ int n_keys;
    __syncthreads();
    if (threadIdx.x == 0) n_keys = 0;
    __syncthreads();
    for (int c = first + threadIdx.x; c < last; c += blockDim.x) {
        const u64 k = cw.kept[c];
        if (k != 0ull) ws.keys[(size_t)b * ws.cap + atomicAdd(&n_keys, 1)] = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) ws.count[b] = n_keys;
}

int g_greedy_rounds = 3;                    // multi-CTA rounds before the per-image finisher (development: balf_debug_set key 6)

static void square_footprint(int r, Footprint* f) {
    f->S = r + 1;
    f->reach = r;
    for (int i = 0; i <= kFpMax; ++i) f->hw[i] = i <= r ? r : -1;
}
// The kept list needs one slot per cell, so in cell mode the key area is re-cut: [B][cells] keys, then the cell structures
// (60 bytes per cell against the 12 bytes per pixel of the key + alive areas: any S >= 3 fits).
static bool cells_layout(const NmsWs& ws, int B, int H, int W, int S, NmsWs* ws2, CellWs* cw) {
    const int cx = cdiv(W, S), cy = cdiv(H, S);
    const size_t cells = (size_t)B * cx * cy;
    const size_t o_alive = align_up(cells * 8, 256), o_cmax = o_alive + align_up(cells * 32, 256);
    const size_t o_kept = o_cmax + align_up(cells * 8, 256);
    const size_t o_dirty = o_kept + align_up(cells * 8, 256), need = o_dirty + align_up(cells * 4, 256);
    const size_t have = align_up(sizeof(u64) * (size_t)H * W * B, 256) + sizeof(float) * (size_t)H * W * B;
    if (S < 3 || S > 16 || need > have) return false;
    char* p = reinterpret_cast<char*>(ws.keys);
    *ws2 = ws;
    ws2->cap = (size_t)cx * cy;
    cw->alive = reinterpret_cast<uint32_t*>(p + o_alive);
    cw->cmax = reinterpret_cast<u64*>(p + o_cmax);
    cw->kept = reinterpret_cast<u64*>(p + o_kept);
    cw->dirty = reinterpret_cast<int*>(p + o_dirty);
    cw->vec_ok = 0;
    cw->cx = cx;
    cw->cy = cy;
    return true;
}
static int run_greedy_cells(const MapView& mv, const NmsWs& ws, const CellWs& cw, const Footprint& f, int B, float thr,
                            cudaStream_t st) {
    const int nc = cw.cx * cw.cy;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = cdiv(nc * B, 8);
    const int grid = want < sms * 8 ? want : sms * 8;                 // 8 CTAs x 8 warps = a full SM of warps
    CellWs cwv = cw;
    cwv.vec_ok = f.S == 16 && mv.left % 4 == 0 && mv.Ws % 4 == 0 && reinterpret_cast<uintptr_t>(mv.score) % 16 == 0;
    BALF_CUDA_OK(cudaMemsetAsync(cw.kept, 0, align_up(sizeof(u64) * (size_t)nc * B, 256) + sizeof(int) * (size_t)nc * B, st));   // kept + dirty
    {
        ProfScope p("nms_greedy_init", st);
        greedy_cells_init_kernel<<<grid, 256, 0, st>>>(mv, cwv, f, thr, B);
    }
    for (int r = 0; r < g_greedy_rounds; ++r) {
        if (r) {
            ProfScope p("nms_greedy_refresh", st);
            greedy_cells_refresh_kernel<<<grid, 256, 0, st>>>(mv, cw, f, B);
        }
        ProfScope p("nms_greedy_decide", st);
        greedy_cells_decide_kernel<<<grid, 256, 0, st>>>(mv, cw, f, ws, B);
    }
    {
        ProfScope p("nms_greedy_finish", st);
        greedy_cells_finish_kernel<<<B, 1024, 0, st>>>(mv, cw, f, ws);
    }
    BALF_COUNT_LAUNCH(2 * g_greedy_rounds + 1);
    BALF_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------ select + sort
// k-th largest value of field(key) among keys satisfying pred -- MSB-first byte-wise radix select.
template <typename Pred, typename Field>
__device__ u64 radix_select(const u64* keys, int n, int kth, int nbytes, Pred pred, Field field, unsigned* hist,
                            unsigned* xch) {
    u64 prefix = 0, mask = 0;
    int remaining = kth;
    for (int shift = (nbytes - 1) * 8; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            u64 k = keys[i];
            if (pred(k)) {
                u64 f = field(k);
                if ((f & mask) == prefix) atomicAdd(&hist[(unsigned)(f >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // bins from the top: lane l owns bins 255 - 8 l .. 248 - 8 l; warp prefix sums find the bin that holds the k-th
            // key (a single thread walking 256 bins was ~2500 cycles of dependent shared-memory loads per digit)
            const int lane = threadIdx.x, hi = 255 - 8 * lane;
            int own = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) own += (int)hist[hi - j];
            int incl = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const unsigned cross = __ballot_sync(0xffffffffu, incl >= remaining);
            const int src = cross ? __ffs(cross) - 1 : 31;
            if (lane == src) {
                int cum = incl - own, bin = hi;
                for (int j = 0; j < 8; ++j, --bin) {
                    const int h = (int)hist[bin];
                    if (cum + h >= remaining || bin == 0) break;
                    cum += h;
                }
                xch[0] = (unsigned)bin;
                xch[1] = (unsigned)(remaining - cum);
            }
        }
        __syncthreads();
        prefix |= (u64)xch[0] << shift;
        mask |= (u64)255 << shift;
        remaining = (int)xch[1];
        __syncthreads();
    }
    return prefix;
}

// soft_argmax_points (test_utils.py:170-215): ps x ps window of the border-masked map zero-padded by
// ps/2 -> normalise by the sum (+1e-6) -> negatives to 1e-6 -> log -> spatial soft-argmax
// (torchgeometry: e = exp(l - max l), centroid = sum(pos * e) / (sum e + 1e-6)) -> minus ps/2.
__global__ void subpixel_kernel(MapView mv, const int32_t* __restrict__ xy, const int32_t* __restrict__ count,
                                int n_max, int ps, float* __restrict__ dxdy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= n_max || (count && i >= count[b])) return;
    const int x = xy[((size_t)b * n_max + i) * 2], y = xy[((size_t)b * n_max + i) * 2 + 1];
    const int half = ps / 2, H = mv.H, W = mv.W;
    auto at = [&](int j) -> float {
        int yy = y - half + j / ps, xx = x - half + j % ps;
        return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? mv.masked(b, yy, xx) : 0.f;
    };
    float sum = 0.f;
    for (int j = 0; j < ps * ps; ++j) sum += at(j);
    const float inv = sum + 1e-6f;
    float mx = kNegInf;
    for (int j = 0; j < ps * ps; ++j) {
        float qv = at(j) / inv;
        if (qv < 0.f) qv = 1e-6f;
        mx = fmaxf(mx, logf(qv));
    }
    float se = 0.f, sx = 0.f, sy = 0.f;
    for (int j = 0; j < ps * ps; ++j) {
        float qv = at(j) / inv;
        if (qv < 0.f) qv = 1e-6f;
        float e = expf(logf(qv) - mx);
        se += e;
        sx += e * (float)(j % ps);
        sy += e * (float)(j / ps);
    }
    const float w = 1.0f / (se + 1e-6f);
    dxdy[((size_t)b * n_max + i) * 2] = sx * w - (float)half;
    dxdy[((size_t)b * n_max + i) * 2 + 1] = sy * w - (float)half;
}

// dense apply_nms output: out = score * (score == window max)
template <int LO, int HI>
__global__ void __launch_bounds__(256) nms_map_kernel(MapView mv, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = Tile<LO, HI, float>;
    const int b = blockIdx.z;
    auto load = [&](int y, int x) -> float {
        return (y >= 0 && y < mv.H && x >= 0 && x < mv.W) ? mv.masked(b, y, x) : kNegInf;
    };
    auto emit = [&](int y, int x, float c, float m) {
        out[((size_t)b * mv.H + y) * mv.W + x] = (c == m) ? c : c * 0.0f;
    };
    run_tile<LO, HI, float>(reinterpret_cast<float*>(smem_raw), blockIdx.y * C::TH, blockIdx.x * C::TW, mv.H, mv.W, load, emit);
}
__global__ void nms_map_generic_kernel(MapView mv, float* __restrict__ out, int lo, int hi) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= mv.W) return;
    float c = mv.masked(b, y, x), m = c;
    int y0 = max(y - lo, 0), y1 = min(y + hi, mv.H - 1), x0 = max(x - lo, 0), x1 = min(x + hi, mv.W - 1);
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) m = fmaxf(m, mv.masked(b, yy, xx));
    out[((size_t)b * mv.H + y) * mv.W + x] = (c == m) ? c : c * 0.0f;
}

// mode 0: windowed (find_index_higher_scores semantics), mode 1: greedy (plain top-k by key)
__global__ void __launch_bounds__(1024) select_sort_kernel(NmsWs ws, int mode, int k, int npow2, int H, int W,
                                                           int32_t* xy, float* out_score, int32_t* out_count, int stage_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* sel = reinterpret_cast<u64*>(smem_raw);          // [npow2]
    u64* stage = sel + npow2;                             // [stage_cap] the image's keys, staged once when they fit: the
                                                          // digit passes of the radix select then run from shared memory
    __shared__ unsigned hist[256];
    __shared__ unsigned xch[2];
    __shared__ int n_sel, c_gt, c_eq;
    const int b = blockIdx.x;
    const u64* keys = ws.keys + (size_t)b * ws.cap;
    const int n = min(ws.count[b], (int)ws.cap);
    if (n > k && n <= stage_cap) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) stage[i] = keys[i];
        keys = stage;                                      // (visible after the barrier below)
    }
    int32_t* oxy = xy + (size_t)b * k * 2;
    float* osc = out_score + (size_t)b * k;

    if (mode == 0 && n == 0) {
        // all-zero NMS map: threshold 0.0 selects the first k raster pixels (test_utils.py:85-93)
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            oxy[2 * i] = i % W;
            oxy[2 * i + 1] = i / W;
            osc[i] = 0.0f;
        }
        if (threadIdx.x == 0) out_count[b] = k;
        return;
    }
    if (threadIdx.x == 0) { n_sel = 0; c_gt = 0; c_eq = 0; }
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) sel[i] = 0ull;
    __syncthreads();

    auto all = [](u64) { return true; };
    if (n <= k) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sel[i] = keys[i];
        if (threadIdx.x == 0) n_sel = n;
    } else if (mode == 1) {
        u64 kth = radix_select(keys, n, k, 8, all, [](u64 q) { return q; }, hist, xch);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            u64 q = keys[i];
            if (q >= kth) sel[atomicAdd(&n_sel, 1)] = q;      // keys are unique -> exactly k
        }
    } else {
        // t = k-th largest score; keep score >= t, and if ties at t overflow k keep the k
        // raster-first of them (even over higher scores) -- quirk (i) of find_index_higher_scores
        u64 t = radix_select(keys, n, k, 4, all, [](u64 q) { return q >> 32; }, hist, xch);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            u64 s = keys[i] >> 32;
            if (s > t) atomicAdd(&c_gt, 1);
            else if (s == t) atomicAdd(&c_eq, 1);
        }
        __syncthreads();
        u64 low = 0;
        if (c_gt + c_eq > k)
            low = radix_select(keys, n, k, 4, [t](u64 q) { return (q >> 32) >= t; },
                               [](u64 q) { return q & 0xFFFFFFFFull; }, hist, xch);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            u64 q = keys[i];
            if ((q >> 32) >= t && (q & 0xFFFFFFFFull) >= low) sel[atomicAdd(&n_sel, 1)] = q;
        }
    }
    __syncthreads();
    // bitonic sort, descending
    if (npow2 == 2048 && blockDim.x == 1024) {
        // Two keys per thread (elements 2t, 2t + 1) in registers: stride 1 is a compare-exchange inside the thread, strides 2-32
        // are 64-bit shuffles inside the warp, and only strides >= 64 (15 of the 66 stages) go through shared memory with a CTA
        // barrier each -- the all-shared-memory version spent two thirds of this kernel's time in its 66 barrier-separated stages.
        const int t = threadIdx.x;
        u64 k0 = sel[2 * t], k1 = sel[2 * t + 1];
        for (int size = 2; size <= 2048; size <<= 1) {
            const bool desc = ((2 * t) & size) == 0;
            int stride = size >> 1;
            if (stride >= 64) {
                sel[2 * t] = k0; sel[2 * t + 1] = k1;
                __syncthreads();
                for (; stride >= 64; stride >>= 1) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool d2 = (lo & size) == 0;
                    const u64 a = sel[lo], c = sel[hi];
                    if ((a < c) == d2) { sel[lo] = c; sel[hi] = a; }
                    __syncthreads();
                }
                k0 = sel[2 * t]; k1 = sel[2 * t + 1];
            }
            for (; stride >= 2; stride >>= 1) {
                const int d = stride >> 1;
                const bool keep_max = ((t & d) == 0) == desc;
                const u64 o0 = __shfl_xor_sync(0xffffffffu, k0, d), o1 = __shfl_xor_sync(0xffffffffu, k1, d);
                k0 = keep_max ? (k0 > o0 ? k0 : o0) : (k0 < o0 ? k0 : o0);
                k1 = keep_max ? (k1 > o1 ? k1 : o1) : (k1 < o1 ? k1 : o1);
            }
            if ((k0 < k1) == desc) { const u64 x = k0; k0 = k1; k1 = x; }
        }
        sel[2 * t] = k0; sel[2 * t + 1] = k1;
        __syncthreads();
    } else {
    for (int size = 2; size <= npow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < npow2 / 2; i += blockDim.x) {
                int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                bool desc = (lo & size) == 0;
                u64 a = sel[lo], c = sel[hi];
                if ((a < c) == desc) { sel[lo] = c; sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    }
    const int m = min(n_sel, k);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        u64 q = sel[i];
        int p = (int)key_raster(q), y = p / W, x = p - y * W;
        oxy[2 * i] = x;
        oxy[2 * i + 1] = y;
        osc[i] = key_score(q);
    }
    if (threadIdx.x == 0) out_count[b] = m;
}

// ------------------------------------------------------------------------------------------ host side
static size_t nms_ws_layout(int B, int H, int W, NmsWs* ws, void* base) {
    size_t off = 0;
    char* p = static_cast<char*>(base);
    size_t cap = (size_t)H * W;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_count = take(sizeof(int32_t) * B), o_flags = take(sizeof(int32_t) * B);
    size_t o_keys = take(sizeof(u64) * cap * B), o_alive = take(sizeof(float) * cap * B);
    if (ws) {
        ws->count = reinterpret_cast<int32_t*>(p + o_count);
        ws->flags = reinterpret_cast<int32_t*>(p + o_flags);
        ws->keys = reinterpret_cast<u64*>(p + o_keys);
        ws->alive = reinterpret_cast<float*>(p + o_alive);
        ws->cap = cap;
    }
    return off;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <int LO, int HI>
static int launch_windowed(const MapView& mv, const NmsWs& ws, int B, cudaStream_t st) {
    using C = Tile<LO, HI, float>;
    size_t smem = C::smem_bytes;
    BALF_CUDA_OK(cudaFuncSetAttribute(windowed_nms_kernel<LO, HI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(mv.W, C::TW), cdiv(mv.H, C::TH), B);
    {
        ProfScope p("nms_windowed", st);
        windowed_nms_kernel<LO, HI><<<grid, 256, smem, st>>>(mv, ws);
    }
    BALF_COUNT_LAUNCH(1);
    return 0;
}

// nms_size 15: the fused two-level kernel
int g_nms_tma = 1;          // development switch (balf_debug_set key 7): 0 = per-thread 128-bit loads for every tile
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static int launch_windowed15(const MapView& mv, const NmsWs& ws, int B, cudaStream_t st) {
    const int vec_ok = (mv.left % 4 == 0) && (mv.Ws % 4 == 0) && (reinterpret_cast<uintptr_t>(mv.score) % 16 == 0);
    const int tiles_x = cdiv(mv.W, kNmsTile), tiles_y = cdiv(mv.H, kNmsTile);
    // 3-D tensor map over the score maps [B, Hs, Ws] fp32, box = one 80 x 80 input window (row pitch, base and box start multiples of 16 bytes)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tma = 0;
    // (the box must start on a 16-byte boundary of its row as well: a crop whose left edge is not a multiple of four pixels
    // faults in the TMA unit -- "illegal instruction" -- so those maps take the per-thread loads)
    if (g_nms_tma && vec_ok && mv.Ws >= kNmsIn && mv.Hs >= kNmsIn) {
        if (EncodeTiledFn enc = encode_tiled_fn()) {
            const cuuint64_t dims[3] = {(cuuint64_t)mv.Ws, (cuuint64_t)mv.Hs, (cuuint64_t)B};
            const cuuint64_t strides[2] = {(cuuint64_t)mv.Ws * 4, (cuuint64_t)mv.Ws * mv.Hs * 4};
            const cuuint32_t box[3] = {(cuuint32_t)kNmsIn, (cuuint32_t)kNmsIn, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            use_tma = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(mv.score), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
    }
    if (use_tma) {
        // persistent CTAs, two window buffers each (50 KB of dynamic shared memory): four CTAs per SM
        static int per_sm = 0, sms = 0;
        const int smem = 2 * kNmsPxBytes;
        if (!per_sm) {
            BALF_CUDA_OK(cudaFuncSetAttribute(nms15_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int dev = 0, occ = 0;
            BALF_CUDA_OK(cudaGetDevice(&dev));
            BALF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            BALF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nms15_tma_kernel, 256, smem));
            BALF_REQUIRE(occ > 0, "internal: nms15_tma_kernel does not fit an SM");
            per_sm = occ;
        }
        const int ntiles = tiles_x * tiles_y * B;
        const int grid = std::min(ntiles, sms * per_sm);
        ProfScope p("nms_windowed", st);
        nms15_tma_kernel<<<grid, 256, smem, st>>>(mv, ws, tmap, tiles_x, tiles_y, ntiles);
    } else {
        dim3 grid(tiles_x, tiles_y, B);
        ProfScope p("nms_windowed", st);
        nms15_kernel<<<grid, 256, 0, st>>>(mv, ws, vec_ok);
    }
    BALF_COUNT_LAUNCH(1);
    return 0;
}

template <int R>
static int launch_greedy(const NmsWs& ws, int B, int H, int W, const Footprint& fp, cudaStream_t st) {
    size_t smem = R > 0 ? Tile<(R > 0 ? R : 1), (R > 0 ? R : 1), u64>::smem_bytes : 0;
    if (smem > 48 * 1024)
        BALF_CUDA_OK(cudaFuncSetAttribute(greedy_nms_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfScope p("nms_greedy_rounds", st);
        greedy_nms_kernel<R><<<B, 256, smem, st>>>(ws, H, W, fp);
    }
    BALF_COUNT_LAUNCH(1);
    return 0;
}

static int check_common(const float* score, int B, int Hs, int Ws, int top, int left, int H, int W, int border,
                        int k, const void* xy, const void* sc, const void* cnt, void* wsp, size_t wsb) {
    BALF_REQUIRE(score && xy && sc && cnt && wsp, "null pointer argument");
    BALF_REQUIRE(B > 0 && H > 0 && W > 0 && k > 0, "B, H, W, k must be positive (B=%d H=%d W=%d k=%d)", B, H, W, k);
    BALF_REQUIRE(top >= 0 && left >= 0 && top + H <= Hs && left + W <= Ws, "crop [%d+%d, %d+%d] exceeds map %dx%d",
                 top, H, left, W, Hs, Ws);
    BALF_REQUIRE(border >= 0, "border must be >= 0");
    BALF_REQUIRE(k <= 16384, "k = %d exceeds the supported maximum of 16384", k);
    BALF_REQUIRE((size_t)H * W < 0x7FFFFFFFull, "map too large");
    BALF_REQUIRE(wsb >= nms_ws_layout(B, H, W, nullptr, nullptr), "workspace too small: %zu < %zu", wsb,
                 nms_ws_layout(B, H, W, nullptr, nullptr));
    return 0;
}

static int run_select(const NmsWs& ws, int mode, int B, int H, int W, int k, int32_t* xy, float* sc, int32_t* cnt,
                      cudaStream_t st) {
    int np2 = next_pow2(k < 2 ? 2 : k);
    const size_t room = (size_t)200 * 1024 - sizeof(u64) * (size_t)np2;
    const int stage_cap = (int)std::min<size_t>(8192, room / sizeof(u64));
    size_t smem = sizeof(u64) * ((size_t)np2 + stage_cap);
    if (smem > 48 * 1024)
        BALF_CUDA_OK(cudaFuncSetAttribute(select_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfScope p("nms_select_sort", st);
        select_sort_kernel<<<B, 1024, smem, st>>>(ws, mode, k, np2, H, W, xy, sc, cnt, stage_cap);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

}  // namespace balf

using namespace balf;

extern "C" size_t balf_nms_workspace_bytes(int B, int H, int W, int k) {
    (void)k;
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return nms_ws_layout(B, H, W, nullptr, nullptr);
}

extern "C" int balf_windowed_nms_topk(const float* score, int B, int Hs, int Ws, int top, int left, int H, int W,
                                      int border, int nms_size, int k, int32_t* xy, float* out_score,
                                      int32_t* count, void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_common(score, B, Hs, Ws, top, left, H, W, border, k, xy, out_score, count, workspace, workspace_bytes)) return e;
    BALF_REQUIRE(nms_size >= 1, "nms_size must be >= 1");
    BALF_REQUIRE((long long)H * W >= k, "score map has fewer than k = %d elements (the reference raises IndexError)", k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NmsWs ws;
    nms_ws_layout(B, H, W, &ws, workspace);
    MapView mv{score, Hs, Ws, top, left, H, W, border};
    BALF_CUDA_OK(cudaMemsetAsync(ws.count, 0, align_up(sizeof(int32_t) * B, 256) + sizeof(int32_t) * B, st));
    const int lo = nms_size / 2, hi = (nms_size - 1) / 2;
    int e = 0;
    if (nms_size == 15) e = launch_windowed15(mv, ws, B, st);
    else if (nms_size == 31) e = launch_windowed<15, 15>(mv, ws, B, st);
    else {
        dim3 grid(cdiv(W, 128), H, B);
        ProfScope p("nms_windowed_generic", st);
        windowed_nms_generic_kernel<<<grid, 128, 0, st>>>(mv, ws, lo, hi);
        BALF_COUNT_LAUNCH(1);
    }
    if (e) return e;
    BALF_LAUNCH_OK();
    return run_select(ws, 0, B, H, W, k, xy, out_score, count, st);
}

extern "C" int balf_greedy_nms_topk(const float* score, int B, int Hs, int Ws, int top, int left, int H, int W,
                                    int border, float thr, int radius, int k, int subpixel_ps, int32_t* xy,
                                    float* out_score, float* dxdy, int32_t* count, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (int e = check_common(score, B, Hs, Ws, top, left, H, W, border, k, xy, out_score, count, workspace, workspace_bytes)) return e;
    BALF_REQUIRE(radius >= 0 && radius <= kFpMax, "radius must be in [0, %d]", kFpMax);
    BALF_REQUIRE(subpixel_ps >= 0 && subpixel_ps <= 15, "sub-pixel patch size must be in [0, 15]");
    BALF_REQUIRE(subpixel_ps == 0 || dxdy != nullptr, "dxdy must be given when sub-pixel refinement is on");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NmsWs ws;
    nms_ws_layout(B, H, W, &ws, workspace);
    MapView mv{score, Hs, Ws, top, left, H, W, border};
    BALF_CUDA_OK(cudaMemsetAsync(ws.count, 0, align_up(sizeof(int32_t) * B, 256) + sizeof(int32_t) * B, st));
    CellWs cw;
    Footprint fp;
    NmsWs wsc;
    square_footprint(radius, &fp);
    if (radius >= 2 && radius <= 15 && g_greedy_impl == 0 && cells_layout(ws, B, H, W, fp.S, &wsc, &cw)) {
        ws = wsc;                                     // kept list: one slot per cell
        if (int e = run_greedy_cells(mv, ws, cw, fp, B, thr, st)) return e;
    } else {
        // radii outside [2, 15] (and the development switch balf_debug_set(5, 1)): whole-image rounds, one CTA per image
        {
            dim3 grid(cdiv(W, 128), H, B);
            ProfScope p("nms_greedy_init", st);
            greedy_init_kernel<<<grid, 128, 0, st>>>(mv, ws, thr);
            BALF_COUNT_LAUNCH(1);
            BALF_LAUNCH_OK();
        }
        int e = radius == 15 ? launch_greedy<15>(ws, B, H, W, fp, st) : launch_greedy<0>(ws, B, H, W, fp, st);
        if (e) return e;
        BALF_LAUNCH_OK();
    }
    if (int e2 = run_select(ws, 1, B, H, W, k, xy, out_score, count, st)) return e2;
    if (subpixel_ps > 0) {
        dim3 grid(cdiv(k, 128), B);
        ProfScope p("nms_subpixel", st);
        subpixel_kernel<<<grid, 128, 0, st>>>(mv, xy, count, k, subpixel_ps, dxdy);
        BALF_COUNT_LAUNCH(1);
        BALF_LAUNCH_OK();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ box_nms
// balf/benchmark_test/repeatability_tools.py:227-255: every pixel with prob >= min_prob is the centre of a size x size box;
// torchvision.ops.nms keeps boxes in descending score order, dropping a box whose IoU with a kept one exceeds `iou`.  All
// boxes have the same size, so "IoU > iou" is a fixed set of pixel offsets: the same greedy NMS with that footprint.  IoU is
// evaluated in float32 exactly like torchvision's kernel: inter / (area_a + area_b - inter) > iou.
__global__ void box_scatter_keys_kernel(const u64* __restrict__ keys, const int32_t* __restrict__ count, size_t cap, int HW,
                                        float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= count[b]) return;
    const u64 k = keys[(size_t)b * cap + i];
    out[(size_t)b * HW + key_raster(k)] = key_score(k);
}
__global__ void box_scatter_xy_kernel(const int32_t* __restrict__ xy, const float* __restrict__ sc, const int32_t* __restrict__ count,
                                      int k, int W, int HW, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= count[b]) return;
    out[(size_t)b * HW + (size_t)xy[((size_t)b * k + i) * 2 + 1] * W + xy[((size_t)b * k + i) * 2]] = sc[(size_t)b * k + i];
}

extern "C" int balf_box_nms_map(const float* prob, int B, int H, int W, float size, float iou, float min_prob, int keep_top_k,
                                float* out, void* workspace, size_t workspace_bytes, void* stream) {
    BALF_REQUIRE(prob && out && workspace, "null pointer argument");
    BALF_REQUIRE(B > 0 && H > 0 && W > 0 && size > 0.f, "bad box_nms arguments");
    BALF_REQUIRE((size_t)H * W < 0x7FFFFFFFull, "map too large");
    BALF_REQUIRE(keep_top_k <= 16384, "keep_top_k = %d exceeds the supported maximum of 16384", keep_top_k);
    const size_t top_bytes = keep_top_k > 0 ? align_up((size_t)B * keep_top_k * 12, 256) : 0;
    BALF_REQUIRE(workspace_bytes >= nms_ws_layout(B, H, W, nullptr, nullptr) + top_bytes, "workspace too small");
    Footprint fp;
    const float area2 = size * size + size * size;
    int diag = -1;
    fp.reach = -1;
    for (int dy = 0; dy <= kFpMax; ++dy) {
        fp.hw[dy] = -1;
        for (int dx = 0; dx <= kFpMax; ++dx) {
            const float inter = fmaxf(size - (float)dy, 0.f) * fmaxf(size - (float)dx, 0.f);
            if (inter / (area2 - inter) > iou) fp.hw[dy] = dx;
        }
        if (fp.hw[dy] >= 0) fp.reach = dy;
        if (fp.hw[dy] >= dy) diag = dy;
    }
    BALF_REQUIRE(size < (float)kFpMax, "box_nms: box size %.2f exceeds the supported maximum of %d", size, kFpMax);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NmsWs ws, wsc;
    nms_ws_layout(B, H, W, &ws, workspace);
    CellWs cw;
    MapView mv{prob, H, W, 0, 0, H, W, 0};
    BALF_CUDA_OK(cudaMemsetAsync(ws.count, 0, align_up(sizeof(int32_t) * B, 256) + sizeof(int32_t) * B, st));
    BALF_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * H * W, st));
    // cells of S = diag + 1 (every two pixels of a cell exclude each other) tile the problem when the footprint stays within
    // one cell of reach; other footprints (e.g. size 6 at iou 0.3: a plus-shaped set) run the whole-image kernel
    fp.S = diag + 1;
    if (fp.reach < 0) {                                   // iou >= 1: nothing is ever suppressed
        fp.reach = 0; fp.hw[0] = 0; fp.S = 1;
    }
    if (diag >= 2 && fp.S <= 16 && fp.reach <= fp.S && fp.hw[0] <= fp.S && g_greedy_impl == 0 &&
        cells_layout(ws, B, H, W, fp.S, &wsc, &cw)) {
        if (int e = run_greedy_cells(mv, wsc, cw, fp, B, min_prob, st)) return e;
    } else {
        wsc = ws;
        dim3 grid(cdiv(W, 128), H, B);
        greedy_init_kernel<<<grid, 128, 0, st>>>(mv, ws, min_prob);
        BALF_COUNT_LAUNCH(1);
        BALF_LAUNCH_OK();
        if (int e = launch_greedy<0>(ws, B, H, W, fp, st)) return e;
        BALF_LAUNCH_OK();
    }
    if (keep_top_k > 0) {
        char* extra = static_cast<char*>(workspace) + nms_ws_layout(B, H, W, nullptr, nullptr);
        int32_t* xy = reinterpret_cast<int32_t*>(extra);
        float* sc = reinterpret_cast<float*>(extra + (size_t)B * keep_top_k * 8);
        int32_t* cnt = wsc.flags;                         // reused as the selected count
        if (int e = run_select(wsc, 1, B, H, W, keep_top_k, xy, sc, cnt, st)) return e;
        box_scatter_xy_kernel<<<dim3(cdiv(keep_top_k, 256), B), 256, 0, st>>>(xy, sc, cnt, keep_top_k, W, H * W, out);
    } else {
        box_scatter_keys_kernel<<<dim3(cdiv((int)wsc.cap, 256), B), 256, 0, st>>>(wsc.keys, wsc.count, wsc.cap, H * W, out);
    }
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_subpixel_refine(const float* score, int B, int Hs, int Ws, int top, int left, int H, int W,
                                    int border, const int32_t* xy, int n, int ps, float* dxdy, void* stream) {
    BALF_REQUIRE(score && xy && dxdy, "null pointer argument");
    BALF_REQUIRE(B > 0 && n >= 0 && ps >= 1 && ps <= 15, "bad sub-pixel arguments (n=%d ps=%d)", n, ps);
    BALF_REQUIRE(top >= 0 && left >= 0 && top + H <= Hs && left + W <= Ws && border >= 0, "bad crop");
    if (n == 0) return 0;
    MapView mv{score, Hs, Ws, top, left, H, W, border};
    dim3 grid(cdiv(n, 128), B);
    subpixel_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(mv, xy, nullptr, n, ps, dxdy);
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

template <int LO, int HI>
static int launch_nms_map(const MapView& mv, float* out, int B, cudaStream_t st) {
    using C = Tile<LO, HI, float>;
    BALF_CUDA_OK(cudaFuncSetAttribute(nms_map_kernel<LO, HI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes));
    dim3 grid(cdiv(mv.W, C::TW), cdiv(mv.H, C::TH), B);
    nms_map_kernel<LO, HI><<<grid, 256, C::smem_bytes, st>>>(mv, out);
    BALF_COUNT_LAUNCH(1);
    return 0;
}

extern "C" int balf_apply_nms_map(const float* score, int B, int H, int W, int border, int nms_size, float* out,
                                  void* stream) {
    BALF_REQUIRE(score && out, "null pointer argument");
    BALF_REQUIRE(B > 0 && H > 0 && W > 0 && border >= 0 && nms_size >= 1, "bad apply_nms arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MapView mv{score, H, W, 0, 0, H, W, border};
    int e = 0;
    if (nms_size == 15) e = launch_nms_map<7, 7>(mv, out, B, st);
    else if (nms_size == 31) e = launch_nms_map<15, 15>(mv, out, B, st);
    else {
        dim3 grid(cdiv(W, 128), H, B);
        nms_map_generic_kernel<<<grid, 128, 0, st>>>(mv, out, nms_size / 2, (nms_size - 1) / 2);
        BALF_COUNT_LAUNCH(1);
    }
    if (e) return e;
    BALF_LAUNCH_OK();
    return 0;
}
