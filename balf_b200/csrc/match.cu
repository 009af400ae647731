// M1: mutual nearest-neighbour matching with Lowe ratio test (SMNN) for sm_100a.
//
// Reference call site: demo/demo_match.py:104-111
//     dists, idxs = K.feature.match_smnn(desc1, desc2, 0.99)
// kornia 0.7.4 is an un-vendored pip dependency (PARITY UNPINNED; restated in oracle/thirdparty.py
// from SURVEY.md appendix B4): D = cdist(d1, d2) in the matmul form sqrt(max(|a|^2 + |b|^2 - 2 a.b, 0));
// each direction keeps rows whose two smallest distances satisfy v0 / v1 <= th; a pair survives when
// it is the nearest neighbour in BOTH directions; output sorted by the first index, distance = max
// of the two ratios.  Fewer than two descriptors on either side -> no match.
//
// Ties are broken towards the LOWER index (canonical rule; torch.topk leaves it unspecified).
// d(i, j) is evaluated with one fixed summation order for both directions, so the row pass over D and
// the row pass over D^T see bit-identical distances.
//
// First version: the 2 * N1 * N2 * 128 flop distance GEMM runs as an FFMA register-tiled kernel, once
// per direction (top-2 needs a full row, so each CTA owns 64 query rows and streams all keys).
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

constexpr int kDim = 128;
constexpr int QT = 64;      // query rows per CTA
constexpr int KT = 64;      // key rows per inner tile

struct Top2 {
    float v0, v1;
    int i0;
    __device__ __forceinline__ void init() { v0 = v1 = __int_as_float(0x7f800000); i0 = 0x7fffffff; }
    // candidates arrive in arbitrary order: order by (value, index)
    __device__ __forceinline__ void push(float v, int i) {
        if (v < v0 || (v == v0 && i < i0)) { v1 = v0; v0 = v; i0 = i; }
        else if (v < v1) v1 = v;
    }
    __device__ __forceinline__ void merge(const Top2& o, float ov1_unused = 0.f) {
        (void)ov1_unused;
        if (o.v0 < v0 || (o.v0 == v0 && o.i0 < i0)) { v1 = fminf(v0, o.v1); v0 = o.v0; i0 = o.i0; }
        else v1 = fminf(v1, o.v0);
    }
};

__global__ void sqnorm_kernel(const float* __restrict__ d, int n, float* __restrict__ out) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(d + (size_t)r * kDim) + lane);
    const float s = warp_sum((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
    if (lane == 0) out[r] = s;
}

// For each query row q of Q [nq,128]: the two smallest distances to the rows of Kd [nk,128] and the
// index of the smallest.  dm (optional) receives the distance matrix, dm[q * ld_q + k * ld_k].
__global__ void __launch_bounds__(256) nn2_kernel(const float* __restrict__ Q, const float* __restrict__ qn, int nq,
                                                  const float* __restrict__ Kd, const float* __restrict__ kn, int nk,
                                                  float* __restrict__ best0, float* __restrict__ best1, int* __restrict__ arg0,
                                                  float* __restrict__ dm, size_t ld_q, size_t ld_k) {
    extern __shared__ __align__(16) unsigned char nn2_smem[];
    float (*qs)[QT + 4] = reinterpret_cast<float (*)[QT + 4]>(nn2_smem);                                   // [k][query]
    float (*ks)[KT + 4] = reinterpret_cast<float (*)[KT + 4]>(nn2_smem + sizeof(float) * kDim * (QT + 4));   // [k][key]
    // the merge scratch aliases the key tile (used only after the last tile has been consumed)
    float (*s_v0)[QT] = reinterpret_cast<float (*)[QT]>(ks);
    float (*s_v1)[QT] = s_v0 + 16;
    int (*s_i0)[QT] = reinterpret_cast<int (*)[QT]>(s_v1 + 16);
    const int q0 = blockIdx.x * QT, tid = threadIdx.x;
    const int tq = tid & 15, tk = tid >> 4;          // thread tile: 4 queries (tq*4..) x 4 keys (tk*4..)
    for (int i = tid; i < QT * (kDim / 4); i += 256) {
        int r = i / (kDim / 4), c4 = i % (kDim / 4);
        float4 v = (q0 + r < nq) ? __ldg(reinterpret_cast<const float4*>(Q + (size_t)(q0 + r) * kDim) + c4) : make_float4(0, 0, 0, 0);
        qs[c4 * 4 + 0][r] = v.x; qs[c4 * 4 + 1][r] = v.y; qs[c4 * 4 + 2][r] = v.z; qs[c4 * 4 + 3][r] = v.w;
    }
    float qnorm[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) qnorm[a] = (q0 + tq * 4 + a < nq) ? __ldg(qn + q0 + tq * 4 + a) : 0.f;
    Top2 best[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) best[a].init();
    for (int k0 = 0; k0 < nk; k0 += KT) {
        __syncthreads();
        for (int i = tid; i < KT * (kDim / 4); i += 256) {
            int r = i / (kDim / 4), c4 = i % (kDim / 4);
            float4 v = (k0 + r < nk) ? __ldg(reinterpret_cast<const float4*>(Kd + (size_t)(k0 + r) * kDim) + c4) : make_float4(0, 0, 0, 0);
            ks[c4 * 4 + 0][r] = v.x; ks[c4 * 4 + 1][r] = v.y; ks[c4 * 4 + 2][r] = v.z; ks[c4 * 4 + 3][r] = v.w;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
        for (int k = 0; k < kDim; ++k) {                      // fixed order k = 0..127 for every (query, key) pair
            const float4 qa = *reinterpret_cast<const float4*>(&qs[k][tq * 4]);
            const float4 kb = *reinterpret_cast<const float4*>(&ks[k][tk * 4]);
            const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(qv[a], kv[b], acc[a][b]);   // a*b commutes: same bits in both directions
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int kk = k0 + tk * 4 + b;
            if (kk >= nk) continue;
            const float knorm = __ldg(kn + kk);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float d2 = fmaf(-2.0f, acc[a][b], qnorm[a] + knorm);
                const float dist = sqrtf(fmaxf(d2, 0.f));
                best[a].push(dist, kk);
                if (dm && q0 + tq * 4 + a < nq) dm[(size_t)(q0 + tq * 4 + a) * ld_q + (size_t)kk * ld_k] = dist;
            }
        }
    }
    // merge the 16 key-slices of every query
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) { s_v0[tk][tq * 4 + a] = best[a].v0; s_v1[tk][tq * 4 + a] = best[a].v1; s_i0[tk][tq * 4 + a] = best[a].i0; }
    __syncthreads();
    if (tid < QT && q0 + tid < nq) {
        Top2 t;
        t.v0 = s_v0[0][tid]; t.v1 = s_v1[0][tid]; t.i0 = s_i0[0][tid];
        for (int s = 1; s < 16; ++s) {
            Top2 o;
            o.v0 = s_v0[s][tid]; o.v1 = s_v1[s][tid]; o.i0 = s_i0[s][tid];
            t.merge(o);
        }
        best0[q0 + tid] = t.v0; best1[q0 + tid] = t.v1; arg0[q0 + tid] = t.i0;
    }
}

// mutual check + ratio test + ordered compaction (single CTA; n1 <= 65536)
// Partial top-2 results of the key splits (tensor-core path) -> one (v0, v1, i0) per row, merged in split order.
__global__ void merge_splits_kernel(float* __restrict__ v0, float* __restrict__ v1, int* __restrict__ i0, int n, int splits) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    Top2 t;
    t.v0 = v0[r]; t.v1 = v1[r]; t.i0 = i0[r];
    for (int s = 1; s < splits; ++s) {
        Top2 o;
        o.v0 = v0[(size_t)s * n + r]; o.v1 = v1[(size_t)s * n + r]; o.i0 = i0[(size_t)s * n + r];
        t.merge(o);
    }
    v0[r] = t.v0; v1[r] = t.v1; i0[r] = t.i0;
}

__global__ void __launch_bounds__(1024) smnn_select_kernel(const float* __restrict__ a0, const float* __restrict__ a1,
                                                           const int* __restrict__ ai, int n1, const float* __restrict__ b0,
                                                           const float* __restrict__ b1, const int* __restrict__ bi, int n2,
                                                           float th, int32_t* __restrict__ ids, float* __restrict__ dist,
                                                           int32_t* __restrict__ count) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = 0; i0 < n1; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        bool ok = false;
        int j = 0;
        float r12 = 0.f, r21 = 0.f;
        if (i < n1 && n1 >= 2 && n2 >= 2) {
            j = ai[i];
            r12 = a0[i] / a1[i];
            if (r12 <= th && j >= 0 && j < n2) {
                r21 = b0[j] / b1[j];
                ok = (r21 <= th) && (bi[j] == i);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (ok) {
            const int slot = off + __popc(m & ((1u << lane) - 1u));
            ids[2 * slot] = i;
            ids[2 * slot + 1] = j;
            dist[slot] = fmaxf(r12, r21);
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += warp_tot[w]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

int match_tc_nn2(bool swap, const float* Q, const float* qn, int nq, const float* Kd, const float* kn, int nk, int splits,
                 float* best0, float* best1, int* arg0, float* dm, size_t ld_q, size_t ld_k, cudaStream_t st);   // match_tc.cu
int g_match_impl = 0;              // debug hook (balf_debug_set key 3): 0 = tensor cores (3xTF32), 1 = fp32 FFMA kernel
constexpr int kMaxSplits = 8;

// key splits of the tensor-core path: enough CTAs to cover the SMs, at least 256 keys per split
static int match_splits(int nq, int nk) {
    int s = 148 / cdiv(nq, 128);
    if (s > cdiv(nk, 256)) s = cdiv(nk, 256);
    return s < 1 ? 1 : s > kMaxSplits ? kMaxSplits : s;
}

}  // namespace balf

using namespace balf;

extern "C" size_t balf_match_workspace_bytes(int n1, int n2) {
    if (n1 <= 0 || n2 <= 0) return 256;
    return (1 + 3 * kMaxSplits) * align_up(sizeof(float) * (size_t)n1, 256) + (1 + 3 * kMaxSplits) * align_up(sizeof(float) * (size_t)n2, 256);
}

extern "C" int balf_match_smnn(const float* d1, int n1, const float* d2, int n2, int dim, float th, int32_t* ids,
                               float* dist, int32_t* count, float* dm_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
    BALF_REQUIRE(ids && dist && count && workspace, "null pointer argument");
    BALF_REQUIRE(dim == kDim, "descriptor dimension must be %d (got %d)", kDim, dim);
    BALF_REQUIRE(n1 >= 0 && n2 >= 0 && n1 <= 65536 && n2 <= 65536, "descriptor counts out of range (%d, %d)", n1, n2);
    BALF_REQUIRE(workspace_bytes >= balf_match_workspace_bytes(n1, n2), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n1 < 2 || n2 < 2) {                       // kornia: fewer than two descriptors on either side -> no match
        BALF_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
        return 0;
    }
    BALF_REQUIRE(d1 && d2, "null descriptor pointer");
    char* p = static_cast<char*>(workspace);
    const size_t s1 = align_up(sizeof(float) * (size_t)n1, 256), s2 = align_up(sizeof(float) * (size_t)n2, 256);
    float* n1sq = reinterpret_cast<float*>(p); p += s1;
    float* a0 = reinterpret_cast<float*>(p); p += kMaxSplits * s1;        // [splits][n1] partial results, merged into split 0
    float* a1 = reinterpret_cast<float*>(p); p += kMaxSplits * s1;
    int* ai = reinterpret_cast<int*>(p); p += kMaxSplits * s1;
    float* n2sq = reinterpret_cast<float*>(p); p += s2;
    float* b0 = reinterpret_cast<float*>(p); p += kMaxSplits * s2;
    float* b1 = reinterpret_cast<float*>(p); p += kMaxSplits * s2;
    int* bi = reinterpret_cast<int*>(p);
    constexpr size_t kNn2Smem = sizeof(float) * kDim * (QT + 4 + KT + 4);
    {
        ProfScope ps("match_sqnorm", st);
        sqnorm_kernel<<<cdiv(n1, 8), 256, 0, st>>>(d1, n1, n1sq);
        sqnorm_kernel<<<cdiv(n2, 8), 256, 0, st>>>(d2, n2, n2sq);
    }
    if (g_match_impl == 0) {
        ProfScope ps("match_nn2", st);
        const int sa = match_splits(n1, n2), sb = match_splits(n2, n1);
        // NOTE: split s of direction 1 writes a0 + s * n1 (dense, not s1-strided): the merge kernel uses the same stride
        if (int e = match_tc_nn2(false, d1, n1sq, n1, d2, n2sq, n2, sa, a0, a1, ai, dm_out, (size_t)n2, 1, st)) return e;
        if (int e = match_tc_nn2(true, d2, n2sq, n2, d1, n1sq, n1, sb, b0, b1, bi, nullptr, 0, 0, st)) return e;
        if (sa > 1) merge_splits_kernel<<<cdiv(n1, 256), 256, 0, st>>>(a0, a1, ai, n1, sa);
        if (sb > 1) merge_splits_kernel<<<cdiv(n2, 256), 256, 0, st>>>(b0, b1, bi, n2, sb);
        BALF_COUNT_LAUNCH(2);
    } else {
        BALF_CUDA_OK(cudaFuncSetAttribute(nn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNn2Smem));
        ProfScope ps("match_nn2", st);
        nn2_kernel<<<cdiv(n1, QT), 256, kNn2Smem, st>>>(d1, n1sq, n1, d2, n2sq, n2, a0, a1, ai, dm_out, (size_t)n2, 1);
        nn2_kernel<<<cdiv(n2, QT), 256, kNn2Smem, st>>>(d2, n2sq, n2, d1, n1sq, n1, b0, b1, bi, nullptr, 0, 0);
    }
    {
        ProfScope ps("match_select", st);
        smnn_select_kernel<<<1, 1024, 0, st>>>(a0, a1, ai, n1, b0, b1, bi, n2, th, ids, dist, count);
    }
    BALF_COUNT_LAUNCH(5);
    BALF_LAUNCH_OK();
    return 0;
}
