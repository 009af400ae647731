// Hardware probe: tcgen05.mma kind::f16 with an MN-major B operand (instruction-descriptor bit 16) in SWIZZLE_NONE layout, M = 64.
// B[k][n] (k = token, n = channel) is stored like an A operand written by row_to_a16: [n / 8][row = k][8 halves] with ROWS rows
// per chunk plane, i.e. 16 bytes = 8 consecutive n, rows 16 bytes apart, 8 rows = one 128-byte core matrix, chunk planes
// ROWS * 16 bytes apart.  Tries both assignments of (LBO, SBO) to (plane stride, 128).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe_mn scripts/umma_probe_mn.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../balf_b200/csrc/umma.cuh"
using namespace balf::umma;

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
// A16 [64 x K] halves (K-major, chunk-major smem), B16 [K x N] halves; mode 0: LBO = plane stride, SBO = 128; mode 1: swapped
__global__ void probe(const __half* A16, const __half* B16, float* D, int N, int K, int ROWS, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __half* sa = reinterpret_cast<__half*>(smem);
    __half* sb = sa + 64 * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * K; i += blockDim.x) { int r = i / K, k = i % K; sa[(k >> 3) * 64 * 8 + r * 8 + (k & 7)] = A16[i]; }
    for (int i = tid; i < K * N; i += blockDim.x) { int k = i / N, n = i % N; sb[((n >> 3) * ROWS + k) * 8 + (n & 7)] = B16[i]; }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);   // B MN-major
        bool acc = false;
        for (int j = 0; j < K / 16; ++j) {
            uint64_t ad = make_desc(smem_u32(sa) + j * 2 * 64 * 16, 64 * 16, 128);
            const uint32_t plane = ROWS * 16;
            uint64_t bd = mode == 0 ? make_desc(smem_u32(sb) + j * 16 * 16, plane, 128) : make_desc(smem_u32(sb) + j * 16 * 16, 128, plane);
            mma_f16(tm, ad, bd, idesc, acc); acc = true;
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        const int lane = tid & 31;
        if (lane < 16) for (int i = 0; i < 16; ++i) D[(size_t)(warp * 16 + lane) * N + c0 + i] = v[i];     // M = 64: rows 16 w + l at lanes 32 w + l
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}
int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode : {0, 1}) for (int N : {32, 64, 128, 256}) for (int ROWS : {64, 128}) {
        const int K = 64;
        std::vector<__half> A(64 * K), B(K * N);
        std::vector<float> D(64 * N), R(64 * N);
        srand(7 + N + ROWS);
        for (auto& x : A) x = __float2half((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : B) x = __float2half((float)rand() / RAND_MAX - 0.5f);
        for (int i = 0; i < 64; ++i) for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)__half2float(A[i * K + k]) * __half2float(B[k * N + j]);
            R[i * N + j] = (float)s;
        }
        __half *dA, *dB; float* dD;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, D.size() * 4);
        size_t smem = (size_t)64 * K * 2 + (size_t)(N / 8) * ROWS * 16 + 1024;
        probe<<<1, 128, smem>>>(dA, dB, dD, N, K, ROWS, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N=%3d ROWS=%3d CUDA ERROR %s\n", mode, N, ROWS, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - R[i]); if (!(d == d)) d = 1e9; err = fmax(err, d); }
        printf("f16 B MN-major none, M=64  (LBO, SBO) = %s  N=%3d ROWS=%3d  max|err| %.3e  %s\n", mode == 0 ? "(plane, 128)" : "(128, plane)", N, ROWS, err, err < 1e-3 ? "PASS" : "FAIL");
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
