// Repeatability metrics and the homography helpers they use, on the GPU.
//
// Reference semantics (restated in oracle/metrics.py, pinned by tests/golden/r2_metrics.npz):
//   balf/benchmark_test/repeatability_tools.py:379-490  compute_repeatability: O(N1 N2) Python loops over circle overlaps
//                                                       (:492-513 intersection_area / union_area), then two greedy
//                                                       assignments in descending overlap order
//   balf/benchmark_test/repeatability_tools.py:516-614  compute_resize_repeatability (SuperPoint-style distance metric)
//   balf/benchmark_test/geometry_tools.py:7-27          create_common_region_masks (cv2.warpPerspective of border-masked
//                                                       ones, >= 0.75, border mask again)
//   balf/benchmark_test/geometry_tools.py:43-64         apply_homography_to_points
//
// All arithmetic is float64 like the reference's Python floats.  The pair loops become one thread per (src, dst) pair; pairs
// whose overlap reaches 1 - overlap_err are appended as (overlap, flat index) candidates, ordered by (overlap descending,
// flat index ascending) -- the stable order of the reference's argsort under the canonical tie rule -- and one thread walks
// them exactly like the reference's loop (the sums of (1 - overlap) are accumulated in the same order, so they are
// bit-identical to the oracle's whenever the overlaps are).
#include <cub/cub.cuh>
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

typedef unsigned long long u64;
constexpr double kPi = 3.141592653589793;
constexpr double kEps64 = 2.220446049250313e-16;    // np.finfo(float).eps
constexpr double kEps32 = 1.1920928955078125e-07;   // np.finfo(np.float32).eps

__device__ __forceinline__ double isect_area(double R, double r, double d) {      // repeatability_tools.py:492-510
    if (d <= fabs(R - r)) { const double m = fmin(R, r); return kPi * (m * m); }
    if (d >= r + R) return 0.0;
    const double r2 = r * r, R2 = R * R, d2 = d * d;
    const double alpha = acos((d2 + r2 - R2) / (2 * d * r));
    const double beta = acos((d2 + R2 - r2) / (2 * d * R));
    return r2 * alpha + R2 * beta - 0.5 * (r2 * sin(2 * alpha) + R2 * sin(2 * beta));
}
__device__ __forceinline__ double union_area(double r, double R, double inter) { return (kPi * (r * r)) + (kPi * (R * R)) - inter; }

// positive doubles order like their bit patterns
__device__ __forceinline__ u64 dkey(double v) { return (u64)__double_as_longlong(v); }

struct RepWs {
    u64* key[2];            // [cap] overlap bits of the candidates of the single-scale (0) / multi-scale (1) matrix
    uint32_t* idx[2];       // [cap] flat index i * n2 + j
    u64* key_b[2];          // sort double buffers
    uint32_t* idx_b[2];
    int* count;             // [2]
    unsigned char* possible;  // [n1]
    unsigned char* visited;   // [n1 + n2]
    void* cub_tmp;
    size_t cub_bytes;
    uint32_t cap;
};

__global__ void rep_pairs_kernel(const double* __restrict__ src, int n1, const double* __restrict__ dst, int n2, double thr,
                                 double eps, double dist_match, double radius, RepWs ws) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n2) return;
    const double rx = src[4 * i], ry = src[4 * i + 1], rr = src[4 * i + 2];
    const double qx = dst[4 * j], qy = dst[4 * j + 1], rd = dst[4 * j + 2];
    const double ddx = rx - qx, ddy = ry - qy;
    const double dist = sqrt((ddx * ddx) + (ddy * ddy));          // ((dx**2) + (dy**2)) ** 0.5: pow(x, 0.5) == sqrt(x) in IEEE
    if (dist <= dist_match) ws.possible[i] = 1;
    if (dist > 4 * radius) return;
    const double f = radius / (fmax(rr, rd) + kEps64);
    double I = isect_area(f * rr, f * rd, dist);
    const double multi = I / (union_area(f * rr, f * rd, I) + eps);
    I = isect_area(radius, radius, dist);
    const double single = I / (union_area(radius, radius, I) + eps);
    const uint32_t flat = (uint32_t)i * (uint32_t)n2 + (uint32_t)j;
    if (single >= thr) {
        const unsigned s = atomicAdd(reinterpret_cast<unsigned*>(ws.count), 1u);
        if (s < ws.cap) { ws.key[0][s] = dkey(single); ws.idx[0][s] = flat; }
    }
    if (multi >= thr) {
        const unsigned s = atomicAdd(reinterpret_cast<unsigned*>(ws.count + 1), 1u);
        if (s < ws.cap) { ws.key[1][s] = dkey(multi); ws.idx[1][s] = flat; }
    }
}

// One thread walks the ordered candidates (repeatability_tools.py:432-470); results -> scalars[8] =
// rep_single, rep_multi, found_single, found_multi, err_single, err_multi, total points, possible matches
__global__ void rep_assign_kernel(RepWs ws, int n1, int n2, double* __restrict__ scalars, int32_t* __restrict__ corr_s,
                                  int32_t* __restrict__ corr_m, int* __restrict__ overflow) {
    __shared__ int possible_sh;
    if (threadIdx.x == 0) possible_sh = 0;
    __syncthreads();
    int loc = 0;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) loc += ws.possible[i] ? 1 : 0;
    atomicAdd(&possible_sh, loc);
    double found[2] = {0, 0}, err[2] = {0, 0};
    for (int m = 0; m < 2; ++m) {
        for (int i = threadIdx.x; i < n1 + n2; i += blockDim.x) ws.visited[i] = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = ws.count[m];
            if (n > (int)ws.cap) *overflow = 1;
            const int nn = n < (int)ws.cap ? n : (int)ws.cap;
            int32_t* corr = m == 0 ? corr_s : corr_m;
            int nf = 0;
            double e = 0.0;
            for (int c = 0; c < nn; ++c) {
                const uint32_t flat = ws.idx[m][c];
                const int y = (int)(flat / (uint32_t)n2), x = (int)(flat - (uint32_t)y * (uint32_t)n2);
                if (ws.visited[y] || ws.visited[n1 + x]) continue;
                const double v = __longlong_as_double((long long)ws.key[m][c]);
                e += (1 - v);
                corr[2 * nf] = x;
                corr[2 * nf + 1] = y;
                ++nf;
                ws.visited[y] = 1;
                ws.visited[n1 + x] = 1;
            }
            found[m] = nf;
            err[m] = nf == 0 ? 0.0 : e / ((double)nf + kEps64);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double points = (double)(n1 < n2 ? n1 : n2);
        scalars[0] = (found[0] / points) * 100.0;
        scalars[1] = (found[1] / points) * 100.0;
        scalars[2] = found[0];
        scalars[3] = found[1];
        scalars[4] = err[0];
        scalars[5] = err[1];
        scalars[6] = points;
        scalars[7] = (double)possible_sh;
    }
}

static size_t rep_layout(int n1, int n2, RepWs* out, unsigned char* base, size_t cub_bytes) {
    const uint32_t cap = 64u * (uint32_t)(n1 > n2 ? n1 : n2) + 1024u;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return base ? base + o : nullptr; };
    RepWs w;
    for (int m = 0; m < 2; ++m) {
        w.key[m] = reinterpret_cast<u64*>(take(sizeof(u64) * cap));
        w.key_b[m] = reinterpret_cast<u64*>(take(sizeof(u64) * cap));
        w.idx[m] = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap));
        w.idx_b[m] = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap));
    }
    w.count = reinterpret_cast<int*>(take(16));
    w.possible = reinterpret_cast<unsigned char*>(take((size_t)n1 + 16));
    w.visited = reinterpret_cast<unsigned char*>(take((size_t)n1 + n2 + 16));
    w.cub_tmp = take(cub_bytes);
    w.cub_bytes = cub_bytes;
    w.cap = cap;
    if (out) *out = w;
    return off;
}

static size_t rep_cub_bytes(uint32_t cap) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const u64*)nullptr, (u64*)nullptr, (int)cap);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, b, (const u64*)nullptr, (u64*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap);
    return a > b ? a : b;
}

// ------------------------------------------------------------------------------------------ distance-threshold repeatability
// kp [n,3] = (row, col, prob).  Step 1: warp (col, row) through the 3x3 matrix M, keep points whose warped (row, col) lies
// inside `shape`; out = (row', col', prob) when `replace`, else the original point (keep_true_keypoints vs the warp +
// filter_keypoints pair of the reference, :530-566).  Compaction keeps the input order.
__global__ void rr_warp_kernel(const double* __restrict__ kp, int n, const double* __restrict__ M, int sh0, int sh1, int replace,
                               double* __restrict__ tmp, unsigned char* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = kp[3 * i], c = kp[3 * i + 1], p = kp[3 * i + 2];
    // homogeneous (c, r, 1) . M^T  -> (x', y', w')
    const double xw = c * M[0] + r * M[1] + M[2], yw = c * M[3] + r * M[4] + M[5], ww = c * M[6] + r * M[7] + M[8];
    const double wc = xw / ww, wr = yw / ww;            // warped (col, row)
    keep[i] = (wr >= 0) && (wr < sh0) && (wc >= 0) && (wc < sh1);
    tmp[3 * i] = replace ? wr : r;
    tmp[3 * i + 1] = replace ? wc : c;
    tmp[3 * i + 2] = p;
}
// select_k_best (:546-551): the k largest probs among the kept points, ties -> the larger input index (the tail of a stable
// ascending sort).  sel[i] = 1 for the selected points.
__global__ void rr_rank_kernel(const double* __restrict__ tmp, const unsigned char* __restrict__ keep, int n, int k,
                               unsigned char* __restrict__ sel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!keep[i]) { sel[i] = 0; return; }
    const double p = tmp[3 * i + 2];
    int after = 0;
    for (int j = 0; j < n; ++j) {
        const double q = tmp[3 * j + 2];
        after += (keep[j] && (q > p || (q == p && j > i))) ? 1 : 0;
    }
    sel[i] = after < k;
}
// one CTA: ordered compaction of the selected points -> pts [count, 2] (row, col) in input order
__global__ void rr_compact_kernel(const double* __restrict__ tmp, const unsigned char* __restrict__ sel, int n,
                                  double* __restrict__ pts, int* __restrict__ count) {
    __shared__ int part[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt, lo = t * per, hi = min(lo + per, n);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += sel[i] ? 1 : 0;
    part[t] = c;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int i = 0; i < nt; ++i) { const int v = part[i]; part[i] = run; run += v; }
        *count = run;
    }
    __syncthreads();
    int o = part[t];
    for (int i = lo; i < hi; ++i)
        if (sel[i]) { pts[2 * o] = tmp[3 * i]; pts[2 * o + 1] = tmp[3 * i + 1]; ++o; }
}
// nearest neighbour distance of every point of a in b (np.linalg.norm over the coordinate axis, then np.min)
__global__ void rr_min_kernel(const double* __restrict__ a, const int* __restrict__ na, const double* __restrict__ b,
                              const int* __restrict__ nb, double* __restrict__ mind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *na) return;
    const double y = a[2 * i], x = a[2 * i + 1];
    double best = __longlong_as_double(0x7ff0000000000000ll);
    const int n = *nb;
    for (int j = 0; j < n; ++j) {
        const double dy = y - b[2 * j], dx = x - b[2 * j + 1];
        best = fmin(best, sqrt(dy * dy + dx * dx));
    }
    mind[i] = best;
}
// out[6] = repeatability, localization_err, N1, N2, count1, count2  (:583-613)
__global__ void rr_final_kernel(const double* __restrict__ min1, const int* __restrict__ n1p, const double* __restrict__ min2,
                                const int* __restrict__ n2p, double thr, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n1 = *n1p, n2 = *n2p;
    int c1 = 0, c2 = 0;
    double s1 = 0.0, s2 = 0.0;
    if (n2 != 0) for (int i = 0; i < n1; ++i) if (min1[i] <= thr) { ++c1; s1 += min1[i]; }
    if (n1 != 0) for (int i = 0; i < n2; ++i) if (min2[i] <= thr) { ++c2; s2 += min2[i]; }
    double rep = 0.0, loc = -1.0;
    if (n1 + n2 > 0) rep = (double)(c1 + c2) / (double)(n1 + n2) * 100.0;
    if (c1 + c2 > 0) {
        loc = 0.0;
        if (n2 != 0) loc += s1 / (double)(c1 + c2);
        if (n1 != 0) loc += s2 / (double)(c1 + c2);
    } else {
        rep = 0.0;
    }
    out[0] = rep; out[1] = loc; out[2] = n1; out[3] = n2; out[4] = c1; out[5] = c2;
}

// ------------------------------------------------------------------------------------------ homography helpers
struct Mat3 { double m[9]; };

// geometry_tools.py:43-64.  The reference inverts / eigen-decomposes 2x2 matrices per point; in closed form
// e0 e1 = det((A (t I) A^T)^-1) = 1 / (t^2 det(A)^2), so the new radius 1 / (e0 e1)^(1/4) = sqrt(t |det A|), t = r^2 + eps32.
__global__ void homography_points_kernel(const double* __restrict__ pts, int n, Mat3 H, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pts[4 * i], y = pts[4 * i + 1], r = pts[4 * i + 2];
    const double* h = H.m;
    const double nx = h[0] * x + h[1] * y + h[2], ny = h[3] * x + h[4] * y + h[5], den = h[6] * x + h[7] * y + h[8];
    const double t = r * r + kEps32;
    const double d2 = den * den;
    const double fxdx = h[0] / den - nx * h[6] / d2, fxdy = h[1] / den - nx * h[7] / d2;
    const double fydx = h[3] / den - ny * h[6] / d2, fydy = h[4] / den - ny * h[7] / d2;
    const double det = fxdx * fydy - fxdy * fydx;
    out[4 * i] = nx / den;
    out[4 * i + 1] = ny / den;
    out[4 * i + 2] = sqrt(t * fabs(det));
    out[4 * i + 3] = pts[4 * i + 3];
}

// cv2.warpPerspective of a border-masked all-ones image (size src_h x src_w, `border` zero pixels all round), INTER_LINEAR
// with OpenCV's 1/32-pixel coordinate grid and constant-zero border, then >= 0.75 and the output's own border mask.
// Mi = inverse of the matrix passed to cv2 (maps output pixels to source positions).  See oracle/metrics.py.
__global__ void common_mask_kernel(Mat3 Mi, int src_h, int src_w, int out_h, int out_w, int border, unsigned char* __restrict__ mask) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out_w) return;
    unsigned char v = 0;
    if (y >= border && y < out_h - border && x >= border && x < out_w - border) {
        const double* m = Mi.m;
        const int bx = out_w >= 64 ? (x / 64) * 64 : 0;
        const double x1 = (double)(x - bx), fy = (double)y, fbx = (double)bx;
        const double X0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], fbx), __dmul_rn(m[1], fy)), m[2]), __dmul_rn(m[0], x1));
        const double Y0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[3], fbx), __dmul_rn(m[4], fy)), m[5]), __dmul_rn(m[3], x1));
        const double W0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[6], fbx), __dmul_rn(m[7], fy)), m[8]), __dmul_rn(m[6], x1));
        const double Wi = W0 != 0.0 ? 32.0 / W0 : 0.0;
        const double lim = 2147483648.0;
        const long long X = llrint(fmin(fmax(__dmul_rn(X0, Wi), -lim), lim - 1)), Y = llrint(fmin(fmax(__dmul_rn(Y0, Wi), -lim), lim - 1));
        const long long sx = X >> 5, sy = Y >> 5;
        const float ax = (float)(X & 31) / 32.0f, ay = (float)(Y & 31) / 32.0f;
        auto tap = [&](long long yy, long long xx) -> double {
            return (yy >= border && yy < src_h - border && xx >= border && xx < src_w - border) ? 1.0 : 0.0;
        };
        const float wx0 = 1.0f - ax, wy0 = 1.0f - ay;
        const double val = tap(sy, sx) * (double)(wx0 * wy0) + tap(sy, sx + 1) * (double)(ax * wy0) +
                           tap(sy + 1, sx) * (double)(wx0 * ay) + tap(sy + 1, sx + 1) * (double)(ax * ay);
        v = val >= 0.75 ? 1 : 0;
    }
    mask[(size_t)y * out_w + x] = v;
}

static bool inv3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return true;
}

}  // namespace balf

using namespace balf;

extern "C" int balf_apply_homography_to_points(const double* pts, int n, const double* h_host, double* out, void* stream) {
    BALF_REQUIRE(n >= 0 && h_host && (n == 0 || (pts && out)), "apply_homography_to_points: bad arguments");
    if (n == 0) return 0;
    Mat3 H;
    for (int i = 0; i < 9; ++i) H.m[i] = h_host[i];
    homography_points_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, H, out);
    BALF_COUNT_LAUNCH(1);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_common_region_masks(const double* h_dst_2_src_host, int src_h, int src_w, int dst_h, int dst_w, int border,
                                         uint8_t* mask_src, uint8_t* mask_dst, void* stream) {
    BALF_REQUIRE(h_dst_2_src_host && mask_src && mask_dst && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && border >= 0,
                 "common_region_masks: bad arguments");
    // mask_src = warp(ones_dst, H): cv2 inverts H, i.e. source positions = H^-1 (x, y); mask_dst = warp(ones_src, inv_h / inv_h[2,2]),
    // whose own inverse is H up to scale
    Mat3 a, b;
    BALF_REQUIRE(inv3(h_dst_2_src_host, a.m), "common_region_masks: singular homography");
    double invn[9];
    for (int i = 0; i < 9; ++i) invn[i] = a.m[i] / a.m[8];
    BALF_REQUIRE(inv3(invn, b.m), "common_region_masks: singular homography");
    cudaStream_t st = (cudaStream_t)stream;
    common_mask_kernel<<<dim3(cdiv(src_w, 128), src_h), 128, 0, st>>>(a, dst_h, dst_w, src_h, src_w, border, mask_src);
    common_mask_kernel<<<dim3(cdiv(dst_w, 128), dst_h), 128, 0, st>>>(b, src_h, src_w, dst_h, dst_w, border, mask_dst);
    BALF_COUNT_LAUNCH(2);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" size_t balf_repeatability_workspace_bytes(int n1, int n2) {
    if (n1 < 0 || n2 < 0) return 0;
    const uint32_t cap = 64u * (uint32_t)(n1 > n2 ? n1 : n2) + 1024u;
    const size_t rep = rep_layout(n1, n2, nullptr, nullptr, rep_cub_bytes(cap));
    const size_t n = (size_t)(n1 > n2 ? n1 : n2) + 16;
    const size_t rr = 2 * align_up(n * 3 * 8, 256) + 4 * align_up(n, 256) + 2 * align_up(n * 2 * 8, 256) + 2 * align_up(n * 8, 256) + 1024;
    return (rep > rr ? rep : rr) + 256;
}

extern "C" int balf_compute_repeatability(const double* src, int n1, const double* dst, int n2, double overlap_err, double eps,
                                           double dist_match_thresh, double radius_size, double* scalars, int32_t* corr_s,
                                           int32_t* corr_m, int32_t* overflow, void* workspace, size_t workspace_bytes, void* stream) {
    BALF_REQUIRE(n1 > 0 && n2 > 0 && src && dst && scalars && corr_s && corr_m && overflow && workspace,
                 "compute_repeatability: needs two non-empty point sets and output buffers");
    BALF_REQUIRE((double)n1 * (double)n2 < 4294967296.0, "compute_repeatability: n1 * n2 must fit 32 bits");
    BALF_REQUIRE(workspace_bytes >= balf_repeatability_workspace_bytes(n1, n2), "compute_repeatability: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t cap = 64u * (uint32_t)(n1 > n2 ? n1 : n2) + 1024u;
    RepWs ws;
    unsigned char* base = reinterpret_cast<unsigned char*>(align_up((size_t)workspace, 256));
    rep_layout(n1, n2, &ws, base, rep_cub_bytes(cap));
    BALF_CUDA_OK(cudaMemsetAsync(ws.count, 0, 16, st));
    BALF_CUDA_OK(cudaMemsetAsync(ws.possible, 0, (size_t)n1, st));
    BALF_CUDA_OK(cudaMemsetAsync(overflow, 0, sizeof(int32_t), st));
    // candidates beyond the count keep the lowest key so that the fixed-size sorts leave them at the end
    for (int m = 0; m < 2; ++m) {
        BALF_CUDA_OK(cudaMemsetAsync(ws.key[m], 0, sizeof(u64) * cap, st));
        BALF_CUDA_OK(cudaMemsetAsync(ws.idx[m], 0xFF, sizeof(uint32_t) * cap, st));
    }
    rep_pairs_kernel<<<dim3(cdiv(n2, 128), n1), 128, 0, st>>>(src, n1, dst, n2, 1.0 - overlap_err, eps, dist_match_thresh, radius_size, ws);
    for (int m = 0; m < 2; ++m) {
        size_t tb = ws.cub_bytes;
        // stable two-pass order: flat index ascending, then overlap descending
        BALF_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, tb, ws.idx[m], ws.idx_b[m], ws.key[m], ws.key_b[m], (int)cap, 0, 32, st));
        tb = ws.cub_bytes;
        BALF_CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(ws.cub_tmp, tb, ws.key_b[m], ws.key[m], ws.idx_b[m], ws.idx[m], (int)cap, 0, 64, st));
    }
    rep_assign_kernel<<<1, 256, 0, st>>>(ws, n1, n2, scalars, corr_s, corr_m, overflow);
    BALF_COUNT_LAUNCH(2);
    BALF_LAUNCH_OK();
    return 0;
}

extern "C" int balf_resize_repeatability(const double* kp, int n1, const double* wkp, int n2, const double* h_host, int src_h,
                                          int src_w, int dst_h, int dst_w, int keep_k, double dist_thresh, double* out6,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    BALF_REQUIRE(n1 >= 0 && n2 >= 0 && h_host && out6 && workspace && keep_k > 0, "resize_repeatability: bad arguments");
    BALF_REQUIRE(workspace_bytes >= balf_repeatability_workspace_bytes(n1, n2), "resize_repeatability: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)(n1 > n2 ? n1 : n2) + 16;
    unsigned char* q = reinterpret_cast<unsigned char*>(align_up((size_t)workspace, 256));
    auto take = [&](size_t bytes) { unsigned char* o = q; q += align_up(bytes, 256); return o; };
    double* tmp_a = reinterpret_cast<double*>(take(n * 3 * 8));
    double* tmp_b = reinterpret_cast<double*>(take(n * 3 * 8));
    unsigned char* keep_a = take(n); unsigned char* keep_b = take(n);
    unsigned char* sel_a = take(n); unsigned char* sel_b = take(n);
    double* pts_a = reinterpret_cast<double*>(take(n * 2 * 8));
    double* pts_b = reinterpret_cast<double*>(take(n * 2 * 8));
    double* min_a = reinterpret_cast<double*>(take(n * 8));
    double* min_b = reinterpret_cast<double*>(take(n * 8));
    int* cnt = reinterpret_cast<int*>(take(64));
    double* dM = reinterpret_cast<double*>(take(2 * 9 * 8));
    double hm[18];
    for (int i = 0; i < 9; ++i) hm[i] = h_host[i];
    BALF_REQUIRE(inv3(h_host, hm + 9), "resize_repeatability: singular homography");
    BALF_CUDA_OK(cudaMemcpyAsync(dM, hm, sizeof(hm), cudaMemcpyHostToDevice, st));
    BALF_CUDA_OK(cudaMemsetAsync(cnt, 0, 64, st));
    // a = keypoints of the first image warped by H into the second (replace), b = keypoints of the second image whose
    // warp by H^-1 falls inside the first (kept as they are)
    if (n1 > 0) {
        rr_warp_kernel<<<cdiv(n1, 128), 128, 0, st>>>(kp, n1, dM, dst_h, dst_w, 1, tmp_a, keep_a);
        rr_rank_kernel<<<cdiv(n1, 128), 128, 0, st>>>(tmp_a, keep_a, n1, keep_k, sel_a);
        rr_compact_kernel<<<1, 1024, 0, st>>>(tmp_a, sel_a, n1, pts_a, cnt);
    }
    if (n2 > 0) {
        rr_warp_kernel<<<cdiv(n2, 128), 128, 0, st>>>(wkp, n2, dM + 9, src_h, src_w, 0, tmp_b, keep_b);
        rr_rank_kernel<<<cdiv(n2, 128), 128, 0, st>>>(tmp_b, keep_b, n2, keep_k, sel_b);
        rr_compact_kernel<<<1, 1024, 0, st>>>(tmp_b, sel_b, n2, pts_b, cnt + 1);
    }
    if (n1 > 0) rr_min_kernel<<<cdiv(n1, 128), 128, 0, st>>>(pts_a, cnt, pts_b, cnt + 1, min_a);
    if (n2 > 0) rr_min_kernel<<<cdiv(n2, 128), 128, 0, st>>>(pts_b, cnt + 1, pts_a, cnt, min_b);
    rr_final_kernel<<<1, 32, 0, st>>>(min_a, cnt, min_b, cnt + 1, dist_thresh, out6);
    BALF_COUNT_LAUNCH(9);
    BALF_LAUNCH_OK();
    return 0;
}
