// Hardware probe for the tcgen05 shared-memory descriptor / instruction descriptor encodings used
// by balf_b200/csrc/umma.cuh.  Runs D[128 x N] = A[128 x K] * B[N x K]^T (tf32) through one CTA for
// several operand layouts and prints the error of each against a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe scripts/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../balf_b200/csrc/umma.cuh"

using namespace balf::umma;

struct Cfg {
    int N, K;
    int a_mode, b_mode;              // 0 chunk-major (K-major, no swizzle), 1 K-major SW128, 2 MN-major chunk-major
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int a_type, b_type;              // descriptor layout_type field (0 none, 2 SW128)
    int a_mn, b_mn;                  // instruction descriptor major bits
    uint32_t a_step, a_group, a_gstride, b_step, b_group, b_gstride;   // descriptor address of MMA j
    int M;                           // 128 or 64
    int a_rows, a_shift;
    int d_lane;                      // lane offset of the accumulator address (M = 64 interleaving)             // physical rows of the A buffer / row offset of logical row 0 (implicit-GEMM shifted views)
};

__device__ int fill_off(int mode, int rows, int K, int r, int k) {   // float offset
    if (mode == 0) return (k >> 2) * rows * 4 + r * 4 + (k & 3);
    if (mode == 1) return (k >> 5) * rows * 32 + r * 32 + ((((k & 31) >> 2) ^ (r & 7)) << 2) + (k & 3);
    return (r >> 2) * K * 4 + k * 4 + (r & 3);                       // mode 2: MN-major, (n/4) chunks of [K][4]
}

__global__ void probe_kernel(const float* A, const float* B, float* D, Cfg c) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    float* sa = reinterpret_cast<float*>(smem);
    float* sb = sa + c.a_rows * c.K;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * c.K; i += blockDim.x) {
        int r = i / c.K, k = i % c.K;
        sa[fill_off(c.a_mode, c.a_rows, c.K, r + c.a_shift, k)] = (r < c.M) ? A[i] : 0.f;
    }
    for (int i = tid; i < c.N * c.K; i += blockDim.x) {
        int r = i / c.K, k = i % c.K;
        sb[fill_off(c.b_mode, c.N, c.K, r, k)] = B[i];
    }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(c.M, c.N, c.a_mn, c.b_mn);
        for (int j = 0; j < c.K / 8; ++j) {
            uint32_t aa = smem_u32(sa) + c.a_shift * 16 + (j / c.a_group) * c.a_gstride + (j % c.a_group) * c.a_step;
            uint32_t ba = smem_u32(sb) + (j / c.b_group) * c.b_gstride + (j % c.b_group) * c.b_step;
            uint64_t ad = make_desc(aa, c.a_lbo, c.a_sbo) | ((uint64_t)c.a_type << 61);
            uint64_t bd = make_desc(ba, c.b_lbo, c.b_sbo) | ((uint64_t)c.b_type << 61);
            mma_tf32(tm + ((uint32_t)c.d_lane << 16), ad, bd, idesc, j > 0);
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < c.N; c0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D[(size_t)tid * c.N + c0 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

static float tf32r(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

int main(int argc, char** argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1;
    std::vector<Cfg> cfgs;
    std::vector<const char*> names;
    auto add = [&](const char* n, Cfg c) { cfgs.push_back(c); names.push_back(n); };
    for (int N : {32, 64, 256}) for (int K : {8, 32, 64, 128}) {
        uint32_t ra = 128 * 16, rb = N * 16;
        add("Kmaj none  LBO=rows*16 SBO=128", Cfg{N, K, 0, 0, ra, 128, rb, 128, 0, 0, 0, 0, 2 * ra, 1 << 20, 0, 2 * rb, 1 << 20, 0, 128, 128, 0, 0});
    }
    for (int N : {32, 128}) for (int K : {8, 64}) for (int sh : {1, 3, 37}) {
        uint32_t ra = 176 * 16, rb = N * 16;
        add("Kmaj none, A view shifted by rows", Cfg{N, K, 0, 0, ra, 128, rb, 128, 0, 0, 0, 0, 2 * ra, 1 << 20, 0, 2 * rb, 1 << 20, 0, 128, 176, sh, 0});
    }
    for (int N : {32, 64, 256}) for (int K : {8, 64, 128}) {
        uint32_t ra = 128 * 16, rb = N * 16;
        add("M=64 Kmaj none", Cfg{N, K, 0, 0, ra, 128, rb, 128, 0, 0, 0, 0, 2 * ra, 1 << 20, 0, 2 * rb, 1 << 20, 0, 64, 128, 0, 0});
        add("Kmaj SW128 SBO=1024", Cfg{N, K, 1, 1, 16, 1024, 16, 1024, 2, 2, 0, 0, 32, 4, 128 * 128, 32, 4, (uint32_t)N * 128, 128, 128, 0, 0});
        add("A Kmaj none, B MNmaj none SBO=K*16 LBO=128", Cfg{N, K, 0, 2, ra, 128, 128, (uint32_t)K * 16, 0, 0, 0, 1, 2 * ra, 1 << 20, 0, 128, 1 << 20, 0, 128, 128, 0, 0});
        add("A Kmaj none, B MNmaj none LBO=K*16 SBO=128 (swapped)", Cfg{N, K, 0, 2, ra, 128, (uint32_t)K * 16, 128, 0, 0, 0, 1, 2 * ra, 1 << 20, 0, 128, 1 << 20, 0, 128, 128, 0, 0});
    }
    for (int N : {32, 128}) for (int K : {8, 64}) {
        uint32_t ra = 128 * 16, rb = N * 16;
        add("M=64 Kmaj none, D at lane 16", Cfg{N, K, 0, 0, ra, 128, rb, 128, 0, 0, 0, 0, 2 * ra, 1 << 20, 0, 2 * rb, 1 << 20, 0, 64, 128, 0, 16});
    }
    if (only == -2) { printf("%zu\n", cfgs.size()); return 0; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, 128 * 256 * 4); cudaMalloc(&dB, 256 * 256 * 4); cudaMalloc(&dD, 128 * 256 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (size_t ci = 0; ci < cfgs.size(); ++ci) {
        if (only >= 0 && (int)ci != only) continue;
        Cfg c = cfgs[ci];
        std::vector<float> A(128 * c.K), B(c.N * c.K), D(128 * c.N), R(128 * c.N);
        srand(1234 + (int)ci);
        for (auto& x : A) x = tf32r((float)rand() / RAND_MAX - 0.5f);
        for (auto& x : B) x = tf32r((float)rand() / RAND_MAX - 0.5f);
        for (int i = 0; i < 128; ++i) for (int j = 0; j < c.N; ++j) {
            double s = 0;
            for (int k = 0; k < c.K; ++k) s += (double)A[i * c.K + k] * B[j * c.K + k];
            R[i * c.N + j] = (float)s;
        }
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, 128 * c.N * 4);
        size_t smem = (size_t)(c.a_rows + c.N) * c.K * 4 + 1024;
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("[%2zu] %-55s N=%3d K=%3d  CUDA ERROR %s\n", ci, names[ci], c.N, c.K, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, err64 = 0;
        int bad_rows = 0;
        for (int i = 0; i < c.M; ++i) {
            double re = 0;
            for (int j = 0; j < c.N; ++j) {
                double d = fabs((double)D[i * c.N + j] - R[i * c.N + j]);
                if (!(d == d)) d = 1e9;
                re = fmax(re, d);
            }
            if (re > 1e-3) ++bad_rows;
            err = fmax(err, re);
        }
        if (c.M == 64) {   // where did the 64 rows land?  compare reference row i against every lane
            int map[4] = {-1, -1, -1, -1};
            for (int probe = 0; probe < 4; ++probe) {
                int i = probe * 16 + 1;
                for (int lane = 0; lane < 128; ++lane) {
                    double re = 0;
                    for (int j = 0; j < c.N; ++j) re = fmax(re, fabs((double)D[lane * c.N + j] - R[i * c.N + j]));
                    if (re < 1e-3) map[probe] = lane;
                }
            }
            printf("[%2zu] %-55s N=%3d K=%3d  rows 1,17,33,49 found at lanes %d %d %d %d\n", ci, names[ci], c.N, c.K, map[0], map[1], map[2], map[3]);
            continue;
        }
        (void)err64;
        printf("[%2zu] %-55s N=%3d K=%3d  max|err| %.3e  bad rows %3d  %s\n", ci, names[ci], c.N, c.K, err, bad_rows, err < 1e-3 ? "PASS" : "FAIL");
    }
    return 0;
}
