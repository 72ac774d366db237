// APA pileup kernels -- /root/reference/hicpeaks/apa.py:11-46.
//
// apa_submatrix (apa.py:11-28): for every anchor (i, j) the dense (2w+1)^2 window of the balanced matrix,
// skipped when it leaves the matrix, holds a NaN or has mean 0, else divided by its mean.
// apa_analysis (apa.py:30-46): per-window mean (again), percentile filter on those means, average of the
// kept windows along axis 0.  The filter selects on rounding noise (every window was just divided by its own
// mean), so the means must be reproduced bit for bit: numpy's float64 add.reduce over a contiguous array is
// the pairwise summation restated below (8 interleaved partial sums per <= 128-element leaf, leaves split at
// multiples of 8), and the axis-0 mean is a plain sequential sum in window order.
//
// Band layout for this path: row-major bal[r][d], d = c - r in [0, num), NaN kept; the matrix is symmetric.
#pragma once
#include "hp_device.cuh"

namespace hp {

constexpr int kApaThreads = 128;
constexpr int kApaMaxLeaves = 256;      // (2w+1)^2 <= 128 * 256

struct ApaPlan {                         // numpy pairwise_sum(n) unrolled by the host
    int n, nleaf, ncomb;
    int leaf_start[kApaMaxLeaves], leaf_len[kApaMaxLeaves];
    short comb_dst[kApaMaxLeaves], comb_src[kApaMaxLeaves];   // res[dst] = res[dst] + res[src], in recursion post-order
};

// sum of a[0..n) in shared memory exactly as numpy adds it; all threads of the block call this
__device__ __forceinline__ double apa_pairwise(const ApaPlan* __restrict__ P, const double* a, double* part, double* res) {
    const int nleaf = P->nleaf;
    for (int t = threadIdx.x; t < nleaf * 8; t += kApaThreads) {
        const int l = t >> 3, j = t & 7;
        const int st = P->leaf_start[l], len = P->leaf_len[l];
        double r = 0.0;
        if (len >= 8) {
            r = a[st + j];
            for (int i = 8; i < len - (len % 8); i += 8) r = __dadd_rn(r, a[st + i + j]);
        }
        part[t] = r;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < nleaf; l += kApaThreads) {
        const int st = P->leaf_start[l], len = P->leaf_len[l];
        double r;
        if (len < 8) {
            r = 0.0;
            for (int i = 0; i < len; ++i) r = __dadd_rn(r, a[st + i]);
        } else {
            const double* p = part + l * 8;
            r = __dadd_rn(__dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3])), __dadd_rn(__dadd_rn(p[4], p[5]), __dadd_rn(p[6], p[7])));
            for (int i = len - (len % 8); i < len; ++i) r = __dadd_rn(r, a[st + i]);
        }
        res[l] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < P->ncomb; ++k) res[P->comb_dst[k]] = __dadd_rn(res[P->comb_dst[k]], res[P->comb_src[k]]);
    __syncthreads();
    const double s = res[0];
    __syncthreads();
    return s;
}

// upload helper: diagonal-major staging [num][n] -> row-major band [n][num], zero where r + d >= n
__global__ void k_apa_transpose(const double* __restrict__ src, double* __restrict__ dst, long long n, int num) {
    __shared__ double t[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long long r = r0 + threadIdx.x;
        const int d = d0 + k;
        t[k][threadIdx.x] = (d < num && r + d < n) ? src[(size_t)d * n + r] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const long long r = r0 + k;
        const int d = d0 + threadIdx.x;
        if (r < n && d < num) dst[(size_t)r * num + d] = t[threadIdx.x][k];
    }
}

// one block per anchor: gather, NaN / zero-mean rejection, normalise, mean of the normalised window
__global__ void __launch_bounds__(kApaThreads) k_apa_windows(const ApaPlan* __restrict__ P, const double* __restrict__ bal, long long n,
                                                             int num, const int* __restrict__ pi, const int* __restrict__ pj, int w,
                                                             double* __restrict__ wins, unsigned char* __restrict__ valid,
                                                             double* __restrict__ mean_arr) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* a = reinterpret_cast<double*>(smem);
    const int side = 2 * w + 1, cells = side * side;
    double* part = a + cells;
    double* res = part + kApaMaxLeaves * 8;
    const long long k = blockIdx.x;
    const long long i = pi[k], j = pj[k];
    if (!(i - w >= 0 && i + w + 1 <= n && j - w >= 0 && j + w + 1 <= n)) {          // apa.py:18
        if (threadIdx.x == 0) { valid[k] = 0; mean_arr[k] = 0.0; }
        return;
    }
    int has_nan = 0;
    for (int t = threadIdx.x; t < cells; t += kApaThreads) {
        const long long r = i - w + t / side, c = j - w + t % side;
        const long long lo = r < c ? r : c, d = r < c ? c - r : r - c;                // symmetric matrix
        const double v = d < num ? bal[lo * num + d] : 0.0;
        a[t] = v;
        has_nan |= (v != v);
    }
    has_nan = __syncthreads_or(has_nan);
    if (has_nan) {                                                                    // apa.py:20-22
        if (threadIdx.x == 0) { valid[k] = 0; mean_arr[k] = 0.0; }
        return;
    }
    const double mean = __ddiv_rn(apa_pairwise(P, a, part, res), (double)cells);
    if (mean == 0.0) {                                                                // apa.py:23-24
        if (threadIdx.x == 0) { valid[k] = 0; mean_arr[k] = 0.0; }
        return;
    }
    double* out = wins + (size_t)k * cells;
    for (int t = threadIdx.x; t < cells; t += kApaThreads) {
        const double v = __ddiv_rn(a[t], mean);                                       // apa.py:26
        a[t] = v;
        out[t] = v;
    }
    __syncthreads();
    const double m2 = __ddiv_rn(apa_pairwise(P, a, part, res), (double)cells);        // apa.py:33
    if (threadIdx.x == 0) { valid[k] = 1; mean_arr[k] = m2; }
}

// means of windows that are already on the device (apa_analysis on caller-supplied windows, apa.py:33)
__global__ void __launch_bounds__(kApaThreads) k_apa_means(const ApaPlan* __restrict__ P, const double* __restrict__ wins, int cells,
                                                           double* __restrict__ mean_arr) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* a = reinterpret_cast<double*>(smem);
    double* part = a + cells;
    double* res = part + kApaMaxLeaves * 8;
    const double* src = wins + (size_t)blockIdx.x * cells;
    for (int t = threadIdx.x; t < cells; t += kApaThreads) a[t] = src[t];
    __syncthreads();
    const double m = __ddiv_rn(apa_pairwise(P, a, part, res), (double)cells);
    if (threadIdx.x == 0) mean_arr[blockIdx.x] = m;
}

// acc[cell] (+)= sequential sum over the selected windows, in order -- the axis-0 sum of apa.py:37
__global__ void k_apa_accumulate(const double* __restrict__ wins, const long long* __restrict__ sel, long long nsel, int cells,
                                 double* __restrict__ acc_io, int init) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells || nsel == 0) return;
    long long s = 0;
    double acc;
    if (init) { acc = wins[(size_t)sel[0] * cells + t]; s = 1; } else acc = acc_io[t];
    for (; s < nsel; ++s) acc = __dadd_rn(acc, wins[(size_t)sel[s] * cells + t]);
    acc_io[t] = acc;
}

}  // namespace hp
