// Hardware probe: tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate) with the chunk-major (K-major SWIZZLE_NONE) layout
// of umma.cuh -- 16-byte K chunks hold 8 halves -- and mixed use with kind::tf32 on one accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe_f16 scripts/umma_probe_f16.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../balf_b200/csrc/umma.cuh"
using namespace balf::umma;

__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);      // A, B = F16 (format 0), D = F32
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
// A16 [128 x K] halves, B16 [N x K] halves (f16 part); A32 [128 x K2] floats, B32 [N x K2] (tf32 part, K2 may be 0)
__global__ void probe(const __half* A16, const __half* B16, const float* A32, const float* B32, float* D, int N, int K, int K2) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __half* sa = reinterpret_cast<__half*>(smem);
    __half* sb = sa + 128 * K;
    float* sa32 = reinterpret_cast<float*>(sb + N * K);
    float* sb32 = sa32 + 128 * K2;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += blockDim.x) { int r = i / K, k = i % K; sa[(k >> 3) * 128 * 8 + r * 8 + (k & 7)] = A16[i]; }
    for (int i = tid; i < N * K; i += blockDim.x) { int r = i / K, k = i % K; sb[(k >> 3) * N * 8 + r * 8 + (k & 7)] = B16[i]; }
    for (int i = tid; i < 128 * K2; i += blockDim.x) { int r = i / K2, k = i % K2; sa32[(k >> 2) * 128 * 4 + r * 4 + (k & 3)] = A32[i]; }
    for (int i = tid; i < N * K2; i += blockDim.x) { int r = i / K2, k = i % K2; sb32[(k >> 2) * N * 4 + r * 4 + (k & 3)] = B32[i]; }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t id16 = make_idesc_f16(128, N), id32 = make_idesc_tf32(128, N);
        bool acc = false;
        for (int j = 0; j < K / 16; ++j) {          // one MMA = K 16 = two 16-byte chunks
            uint64_t ad = make_desc(smem_u32(sa) + j * 2 * 128 * 16, 128 * 16, 128);
            uint64_t bd = make_desc(smem_u32(sb) + j * 2 * N * 16, N * 16, 128);
            mma_f16(tm, ad, bd, id16, acc); acc = true;
        }
        for (int j = 0; j < K2 / 8; ++j) {
            uint64_t ad = make_desc(smem_u32(sa32) + j * 2 * 128 * 16, 128 * 16, 128);
            uint64_t bd = make_desc(smem_u32(sb32) + j * 2 * N * 16, N * 16, 128);
            mma_tf32(tm, ad, bd, id32, acc); acc = true;
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}
static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }
int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {32, 64, 128}) for (int K : {16, 32, 64, 128}) for (int K2 : {0, 32}) {
        std::vector<__half> A(128 * K), B(N * K);
        std::vector<float> A2(128 * (K2 ? K2 : 1)), B2(N * (K2 ? K2 : 1)), D(128 * N), R(128 * N);
        srand(7 + N + K + K2);
        for (auto& x : A) x = __float2half((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : B) x = __float2half((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : A2) x = tf32r((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : B2) x = tf32r((float)rand() / RAND_MAX - 0.5f);
        for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)__half2float(A[i * K + k]) * __half2float(B[j * K + k]);
            for (int k = 0; k < K2; ++k) s += (double)A2[i * K2 + k] * B2[j * K2 + k];
            R[i * N + j] = (float)s;
        }
        __half *dA, *dB; float *dA2, *dB2, *dD;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dA2, A2.size() * 4); cudaMalloc(&dB2, B2.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dA2, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, D.size() * 4);
        size_t smem = (size_t)(128 + N) * K * 2 + (size_t)(128 + N) * K2 * 4 + 1024;
        probe<<<1, 128, smem>>>(dA, dB, dA2, dB2, dD, N, K, K2);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%3d K=%3d K2=%2d CUDA ERROR %s\n", N, K, K2, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - R[i]); if (!(d == d)) d = 1e9; err = fmax(err, d); }
        printf("f16 chunk-major K-major none  N=%3d K=%3d (+ tf32 K=%2d on the same accumulator)  max|err| %.3e  %s\n", N, K, K2, err, err < 1e-3 ? "PASS" : "FAIL");
        cudaFree(dA); cudaFree(dB); cudaFree(dA2); cudaFree(dB2); cudaFree(dD);
    }
    return 0;
}
