// M1 on the tensor cores: the distance GEMM of SMNN matching as tcgen05.mma tiles with fp32-class accuracy.
//
// Reference call site: demo/demo_match.py:104-111 (kornia match_smnn; PARITY UNPINNED, see match.cu / oracle/thirdparty.py).
// D = Q K^T for 128 query rows x 64 key rows per tile, K = 128.  A single tf32 product (10-bit mantissas) would move
// distances by ~1e-4 and with them nearest-neighbour decisions, so every operand is split x = hi + lo (both exactly
// representable in tf32) and three MMAs accumulate hi*hi + hi*lo + lo*hi in fp32 ("3xTF32": |err| ~ 1e-6 on unit-norm
// descriptors, measured against torch.cdist).  The order of the two cross terms is swapped for the second direction
// (queries = d2) so that d(i, j) sees the same products in the same order both ways: bit-identical distances, which the
// mutual check relies on.  One thread per query row (TMEM lane = row): distance, sqrt and the running top-2 are
// thread-local in the tcgen05.ld epilogue, which runs while the next key tile's MMAs execute (two accumulators).
// Keys are split over blockIdx.y; partial top-2 results are merged in a fixed order by the select kernel.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/balf_b200.h"

namespace balf {
using namespace umma;

constexpr int kMq = 128, kMk = 64, kMd = 128;
constexpr int kKStr = kMk * 4 + 4;                       // floats per 16-byte K-chunk of the key operand (+16 B pad: the
                                                         // transposing stores of a warp then hit eight different bank groups)
constexpr size_t kMatchTcSmem = sizeof(float) * (2 * kMq * kMd + 2 * (kMd / 4) * kKStr + 2 * kMk) + 64;

struct Top2T {
    float v0, v1;
    int i0;
    __device__ __forceinline__ void init() { v0 = v1 = __int_as_float(0x7f800000); i0 = 0x7fffffff; }
    __device__ __forceinline__ void push(float v, int i) {
        if (v < v0 || (v == v0 && i < i0)) { v1 = v0; v0 = v; i0 = i; }
        else if (v < v1) v1 = v;
    }
};

__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
    hi = make_float4(to_tf32_exact(v.x), to_tf32_exact(v.y), to_tf32_exact(v.z), to_tf32_exact(v.w));
    lo = make_float4(to_tf32_exact(v.x - hi.x), to_tf32_exact(v.y - hi.y), to_tf32_exact(v.z - hi.z), to_tf32_exact(v.w - hi.w));
}

// partial results: best0 / best1 / arg0 [gridDim.y][nq]
template <bool SWAP>
__global__ void __launch_bounds__(kMq, 1)
nn2_tc_kernel(const float* __restrict__ Q, const float* __restrict__ qn, int nq, const float* __restrict__ Kd,
              const float* __restrict__ kn, int nk, int keys_per_split, float* __restrict__ best0, float* __restrict__ best1,
              int* __restrict__ arg0, float* __restrict__ dm, size_t ld_q, size_t ld_k) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* qhi = reinterpret_cast<float*>(smem);
    float* qlo = qhi + kMq * kMd;
    float* khi = qlo + kMq * kMd;
    float* klo = khi + (kMd / 4) * kKStr;
    float* kns = klo + (kMd / 4) * kKStr;                 // [2][64] squared norms of the key tile
    uint64_t* done = reinterpret_cast<uint64_t*>(kns + 2 * kMk);
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 2);
    const int tid = threadIdx.x, q = blockIdx.x * kMq + tid;
    if (tid < 32) tmem_alloc(slot, 128);
    if (tid == 0) { mbar_init(&done[0], 1); mbar_init(&done[1], 1); mbar_fence_init(); }
    // query rows -> hi / lo operands (chunk-major, one thread per row)
#pragma unroll 4
    for (int c4 = 0; c4 < kMd / 4; ++c4) {
        const float4 v = q < nq ? __ldg(reinterpret_cast<const float4*>(Q + (size_t)q * kMd) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hi, lo;
        split_tf32(v, hi, lo);
        *reinterpret_cast<float4*>(qhi + ((size_t)c4 * kMq + tid) * 4) = hi;
        *reinterpret_cast<float4*>(qlo + ((size_t)c4 * kMq + tid) * 4) = lo;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *slot;
    const uint32_t lane_base = tm + ((uint32_t)(tid & ~31) << 16);
    const bool w0 = warp0_uniform();
    const float qnorm = q < nq ? __ldg(qn + q) : 0.f;
    const int k_begin = blockIdx.y * keys_per_split, k_end = min(nk, k_begin + keys_per_split);
    const int ntile = k_end > k_begin ? (k_end - k_begin + kMk - 1) / kMk : 0;
    const uint64_t ah = make_desc(smem_u32(qhi), kMq * 16, 128), al = make_desc(smem_u32(qlo), kMq * 16, 128);
    const uint64_t bh = make_desc(smem_u32(khi), kKStr * 4, 128), bl = make_desc(smem_u32(klo), kKStr * 4, 128);
    constexpr uint32_t idesc = make_idesc_tf32(kMq, kMk);
    Top2T best;
    best.init();
    for (int kt = 0; kt <= ntile; ++kt) {
        if (kt >= 1) {                                     // MMAs of tile kt-1 done: its accumulator is ready, the key operand is free
            if (tid == 0) mbar_wait(&done[(kt - 1) & 1], ((kt - 1) >> 1) & 1);
            __syncthreads();
            fence_after_sync();
        }
        if (kt < ntile) {
            const int k0 = k_begin + kt * kMk;
            // 64 key rows x 32 chunks: a warp reads 512 contiguous bytes of one row and scatters them over the chunk planes
#pragma unroll 4
            for (int i = 0; i < (kMk * kMd / 4) / kMq; ++i) {
                const int idx = tid + kMq * i, r = idx >> 5, c4 = idx & 31;
                const float4 v = k0 + r < k_end ? __ldg(reinterpret_cast<const float4*>(Kd + (size_t)(k0 + r) * kMd) + c4)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 hi, lo;
                split_tf32(v, hi, lo);
                *reinterpret_cast<float4*>(khi + (size_t)c4 * kKStr + r * 4) = hi;
                *reinterpret_cast<float4*>(klo + (size_t)c4 * kKStr + r * 4) = lo;
            }
            if (tid < kMk) kns[(kt & 1) * kMk + tid] = k0 + tid < k_end ? __ldg(kn + k0 + tid) : 0.f;
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
            if (w0 && elect_one()) {
                const uint32_t d = tm + (uint32_t)(kt & 1) * kMk;
#pragma unroll
                for (uint32_t k8 = 0; k8 < kMd / 8; ++k8) {
                    const uint64_t ao = (k8 * 2u * (kMq * 16u)) >> 4, bo = (k8 * 2u * (kKStr * 4u)) >> 4;
                    mma_tf32(d, ah + ao, bh + bo, idesc, k8 > 0);
                    if (!SWAP) { mma_tf32(d, ah + ao, bl + bo, idesc, true); mma_tf32(d, al + ao, bh + bo, idesc, true); }
                    else       { mma_tf32(d, al + ao, bh + bo, idesc, true); mma_tf32(d, ah + ao, bl + bo, idesc, true); }
                }
                commit(&done[kt & 1]);
            }
        }
        if (kt >= 1) {                                     // epilogue of tile kt-1, under the MMAs of tile kt
            const int b = (kt - 1) & 1, k0 = k_begin + (kt - 1) * kMk;
            float acc[kMk];
            tmem_ld32(lane_base + b * kMk, *reinterpret_cast<float (*)[32]>(&acc[0]));
            tmem_ld32(lane_base + b * kMk + 32, *reinterpret_cast<float (*)[32]>(&acc[32]));
            tmem_ld_wait();
            fence_before_sync();
#pragma unroll
            for (int j = 0; j < kMk; ++j) {                // fully unrolled: acc[] must stay in registers
                const int kk = k0 + j;
                if (kk < k_end) {
                    const float d2 = fmaf(-2.0f, acc[j], qnorm + kns[b * kMk + j]);
                    const float dist = sqrtf(fmaxf(d2, 0.f));
                    best.push(dist, kk);
                    if (dm && q < nq) dm[(size_t)q * ld_q + (size_t)kk * ld_k] = dist;
                }
            }
        }
    }
    if (q < nq) {
        const size_t o = (size_t)blockIdx.y * nq + q;
        best0[o] = best.v0; best1[o] = best.v1; arg0[o] = best.i0;
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 128);
}

int match_tc_nn2(bool swap, const float* Q, const float* qn, int nq, const float* Kd, const float* kn, int nk, int splits,
                 float* best0, float* best1, int* arg0, float* dm, size_t ld_q, size_t ld_k, cudaStream_t st) {
    const int per = cdiv(cdiv(nk, splits), kMk) * kMk;
    dim3 grid(cdiv(nq, kMq), splits);
    if (swap) {
        BALF_CUDA_OK(cudaFuncSetAttribute(nn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchTcSmem));
        nn2_tc_kernel<true><<<grid, kMq, kMatchTcSmem, st>>>(Q, qn, nq, Kd, kn, nk, per, best0, best1, arg0, dm, ld_q, ld_k);
    } else {
        BALF_CUDA_OK(cudaFuncSetAttribute(nn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchTcSmem));
        nn2_tc_kernel<false><<<grid, kMq, kMatchTcSmem, st>>>(Q, qn, nq, Kd, kn, nk, per, best0, best1, arg0, dm, ld_q, ld_k);
    }
    return 0;
}

}  // namespace balf
