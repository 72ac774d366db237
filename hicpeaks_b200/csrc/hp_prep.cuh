// K0 -- the worker's input preparation on the device (/root/reference/scripts/pyHICCUPS:143-166).
//
// From raw counts and the balancing weights it builds what the reference's worker hands to hiccups():
//   balanced diagonal d   = (w[r] * w[r + d]) * count where a count is stored (cooler's `balance=` transform,
//                           api.py: mat.data = bias1[mat.row] * bias2[mat.col] * mat.data, evaluated left to right),
//                           0 elsewhere, NaN -> 0                                             (:143, :153-157)
//   IR[d]                 = mean of the non-NaN entries of that diagonal, zeros included     (:154-156)
//   biases                = 1 / w, 0 where w is 0 or NaN                                     (:163-166)
// so that only 4 bytes per band pixel cross PCIe instead of 12.  `ndarray.mean()` is numpy's pairwise
// summation (see hp_apa.cuh) over the COMPACTED array of non-NaN entries; the kernel compacts each
// diagonal and walks the same recursion, so IR is bit-identical to the reference's.
#pragma once
#include "hp_device.cuh"

namespace hp {

constexpr int kPrepThreads = 256;

// Count upload: the host narrows every diagonal to the smallest of u8 / u16 / i32 that holds its values
// (hp_hostpack.cpp) so that ~1 byte per band pixel crosses PCIe; k_prep_band / k_unpack_quad read that packed
// form directly (4 bins per thread and load) and write the quad-interleaved planes.
struct PackedDiag {
    unsigned long long off;      // byte offset of the diagonal in the packed buffer (16-byte aligned)
    unsigned int esize;          // 1, 2 or 4
    unsigned int len;            // n - d
};

// four consecutive counts of a narrowed diagonal, starting at bin rb (a multiple of 4); bins >= len read as 0
__device__ __forceinline__ int4 packed_quad(const unsigned char* __restrict__ packed, const PackedDiag& t, int rb) {
    int4 v = make_int4(0, 0, 0, 0);
    if (rb < (int)t.len) {
        const unsigned char* src = packed + t.off;
        if (t.esize == 1) {
            const uchar4 u = *reinterpret_cast<const uchar4*>(src + rb);
            v = make_int4(u.x, u.y, u.z, u.w);
        } else if (t.esize == 2) {
            const ushort4 u = *reinterpret_cast<const ushort4*>(src + (size_t)rb * 2);
            v = make_int4(u.x, u.y, u.z, u.w);
        } else {
            v = *reinterpret_cast<const int4*>(src + (size_t)rb * 4);
        }
        if (rb + 1 >= (int)t.len) v.y = 0;           // the last quad of a diagonal may hold staging bytes past its end
        if (rb + 2 >= (int)t.len) v.z = 0;
        if (rb + 3 >= (int)t.len) v.w = 0;
    }
    return v;
}

// raw-count planes below bal_first (no balanced values there): narrowed diagonal -> quad-interleaved int32 plane
__global__ void __launch_bounds__(256) k_unpack_quad(const unsigned char* __restrict__ packed, const PackedDiag* __restrict__ tab,
                                                     int* __restrict__ raw, int pitch) {
    const int d = blockIdx.y;
    const int rb = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (rb >= pitch) return;
    const int4 v = packed_quad(packed, tab[d], rb);
    const size_t q = qidx(d, rb, pitch);                // rb & 3 == 0: the four bins sit at the same slot of the 4 sub-planes
    const size_t sp = (size_t)(pitch >> 2);
    raw[q] = v.x; raw[q + sp] = v.y; raw[q + 2 * sp] = v.z; raw[q + 3 * sp] = v.w;
}

// One CTA per diagonal d in [bal_first, num): reads the narrowed counts, writes the quad-interleaved raw and balanced
// planes, the row flags, the NaN-compacted balanced values (comp: scratch [num][pitch] doubles) and IR[d].
// Compaction: rounds of 2 x 1024 bins, 4 consecutive bins per thread and quad; warp scans + double-buffered warp totals
// give every thread its output offset with one block barrier per round.
// The summation tree has `depth` levels below the root and nslot = 2^(depth + 1) heap slots: in shared memory when
// it fits (use_smem), otherwise in the scratch arrays tree / tval / tlist, [num][nslot] each.
__global__ void __launch_bounds__(kPrepThreads) k_prep_band(const unsigned char* __restrict__ packed, const PackedDiag* __restrict__ tab,
                                                            const double* __restrict__ w, int n, int num,
                                                            int pitch, int bal_first, int* __restrict__ raw, double* __restrict__ bal,
                                                            unsigned int* __restrict__ rownz,
                                                            double* __restrict__ ir, double* __restrict__ comp, int2* __restrict__ tree,
                                                            double* __restrict__ tval, int* __restrict__ tlist, int depth, int nslot,
                                                            int use_smem, unsigned int* __restrict__ domain_bad) {
    extern __shared__ __align__(16) unsigned char prep_smem[];     // use_smem: the summation tree (nodes, values, leaf list)
    constexpr int kWarps = kPrepThreads / 32;
    constexpr int kU = 2;                               // quads per thread and round: 2 x 1024 bins per round in flight
    __shared__ int sh_scan[2][kU][kWarps];              // warp totals, double-buffered: one barrier per round
    __shared__ int sh_base, sh_nleaf;
    const int d = bal_first + blockIdx.x;
    const int len = n - d;
    const PackedDiag pd = tab[d];
    double* cp = comp + (size_t)blockIdx.x * pitch;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t sp = (size_t)(pitch >> 2);
    int base = 0;                                       // kept values before this round (same in every thread)
    // ---- balanced values, planes, row flags, NaN compaction (order preserved) ----------------------------------
    for (int r0 = 0, it = 0; r0 < pitch; r0 += kU * kPrepThreads * 4, ++it) {
        double v[kU][4];
        int cc[kU][4], nk[kU], inc[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int rb = r0 + u * kPrepThreads * 4 + threadIdx.x * 4;
            const int4 c4 = packed_quad(packed, pd, rb);                         // bins >= len read as 0
            cc[u][0] = c4.x; cc[u][1] = c4.y; cc[u][2] = c4.z; cc[u][3] = c4.w;
            nk[u] = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = rb + k;
                v[u][k] = 0.0;
                if (r < len) {                          // weights loaded whether or not a count is stored: no load waits on another
                    // balanced value where a count is stored: cooler's `bias1[row] * bias2[col] * data`, left to right (:143)
                    const double p = __dmul_rn(__dmul_rn(w[r], w[r + d]), (double)cc[u][k]);
                    if (cc[u][k] != 0) v[u][k] = p;
                    nk[u] += (v[u][k] == v[u][k]);
                }
            }
            int x = nk[u];                              // inclusive scan over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += t;
            }
            inc[u] = x;
            if (lane == 31) sh_scan[it & 1][u][wid] = x;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            int before = 0, total = 0;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) {
                const int t = sh_scan[it & 1][u][k];
                total += t;
                if (k < wid) before += t;
            }
            int pre = base + before + inc[u] - nk[u];
            base += total;
            const int rb = r0 + u * kPrepThreads * 4 + threadIdx.x * 4;
            if (rb < pitch) {
                const size_t q = qidx(d, rb, pitch);    // rb & 3 == 0: the four bins sit at the same slot of the 4 sub-planes
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = rb + k;
                    const bool isn = v[u][k] != v[u][k];
                    if (r < len && !isn) cp[pre++] = v[u][k];
                    const double out = isn ? 0.0 : v[u][k];
                    raw[q + k * sp] = cc[u][k];
                    bal[q + k * sp] = out;
                    if (out != 0.0) rownz[r] = 1u;
                    if (fast_domain_bad(out)) *domain_bad = 1u;
                }
            }
        }
    }
    if (threadIdx.x == 0) sh_base = base;
    __syncthreads();
    const int m = sh_base;
    // ---- numpy pairwise_sum(cp[0..m)) -------------------------------------------------------------------
    // numpy's recursion: a block of n <= 128 values is summed with 8 interleaved accumulators; a longer block is
    // split at n2 = n / 2 - (n / 2) % 8 and its value is value(left) + value(right).  The value of a node depends on
    // its two children only, so the tree is built level by level (heap numbering: children of node k are 2k + 1,
    // 2k + 2), all leaves are summed in parallel, and the levels are folded bottom-up -- the same additions in
    // the same association as the sequential recursion, without one thread walking it.
    int2* node = use_smem ? reinterpret_cast<int2*>(prep_smem) : tree + (size_t)blockIdx.x * nslot;          // (start, n); n < 0: absent
    double* val = use_smem ? reinterpret_cast<double*>(prep_smem + (size_t)nslot * 8) : tval + (size_t)blockIdx.x * nslot;
    int* lst = use_smem ? reinterpret_cast<int*>(prep_smem + (size_t)nslot * 16) : tlist + (size_t)blockIdx.x * nslot;
    if (threadIdx.x == 0) { node[0] = make_int2(0, m); sh_nleaf = 0; }
    __syncthreads();
    for (int l = 0; l < depth; ++l) {
        const int first = (1 << l) - 1;
        for (int t = threadIdx.x; t < (1 << l); t += kPrepThreads) {
            const int2 nd = node[first + t];
            int2 lc = make_int2(0, -1), rc = make_int2(0, -1);
            if (nd.y > 128) {
                int n2 = nd.y / 2;
                n2 -= n2 % 8;
                lc = make_int2(nd.x, n2);
                rc = make_int2(nd.x + n2, nd.y - n2);
            }
            node[2 * (first + t) + 1] = lc;
            node[2 * (first + t) + 2] = rc;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < nslot - 1; k += kPrepThreads) {
        const int ny = node[k].y;
        if (ny >= 0 && ny <= 128) lst[atomicAdd(&sh_nleaf, 1)] = k;
    }
    __syncthreads();
    const int nleaf = sh_nleaf;
    // a leaf: 8 lanes hold numpy's r[0..7]; lane 0 folds them and adds the n % 8 trailing values
    for (int t = threadIdx.x; t < nleaf * 8; t += kPrepThreads) {
        const int k = lst[t >> 3];
        const int2 L = node[k];
        const int j = t & 7;
        double r = 0.0;
        if (L.y >= 8) {
            // a leaf holds <= 128 values: this lane's <= 16 are fetched together (the adds are a dependent chain, the
            // loads are not), then added in numpy's order
            const int nfull = L.y - (L.y % 8);
            double x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = (8 * q < nfull) ? cp[L.x + 8 * q + j] : 0.0;
            r = x[0];
#pragma unroll
            for (int q = 1; q < 16; ++q)
                if (8 * q < nfull) r = __dadd_rn(r, x[q]);
        }
        const unsigned grp = 0xffu << ((threadIdx.x & 31) & ~7);
        const int src = (threadIdx.x & 31) & ~7;
        const double r1 = __shfl_sync(grp, r, src + 1), r2 = __shfl_sync(grp, r, src + 2), r3 = __shfl_sync(grp, r, src + 3);
        const double r4 = __shfl_sync(grp, r, src + 4), r5 = __shfl_sync(grp, r, src + 5), r6 = __shfl_sync(grp, r, src + 6);
        const double r7 = __shfl_sync(grp, r, src + 7);
        if (j == 0) {
            double v;
            if (L.y < 8) {
                v = 0.0;
                for (int i = 0; i < L.y; ++i) v = __dadd_rn(v, cp[L.x + i]);
            } else {
                v = __dadd_rn(__dadd_rn(__dadd_rn(r, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
                for (int i = L.y - (L.y % 8); i < L.y; ++i) v = __dadd_rn(v, cp[L.x + i]);
            }
            val[k] = v;
        }
    }
    __syncthreads();
    for (int l = depth - 1; l >= 0; --l) {
        const int first = (1 << l) - 1;
        for (int t = threadIdx.x; t < (1 << l); t += kPrepThreads) {
            const int k = first + t;
            if (node[k].y > 128) val[k] = __dadd_rn(val[2 * k + 1], val[2 * k + 2]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ir[d] = __ddiv_rn(m ? val[0] : 0.0, (double)m);     // mean of an empty slice is NaN, as numpy's
}

// biases = 1 / w (0 where w is 0 or NaN) -- pyHICCUPS:163-166
__global__ void k_prep_bias(const double* __restrict__ w, double* __restrict__ b1, double* __restrict__ b2, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double v = w[r];
    const double b = (v == 0.0 || v != v) ? 0.0 : __ddiv_rn(1.0, v);
    b1[r] = b;
    b2[r] = b;
}

// planes below bal_first hold no balanced values
__global__ void k_zero_planes(double* __restrict__ bal, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) bal[i] = 0.0;
}

}  // namespace hp
