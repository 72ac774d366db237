// K0 -- the worker's input preparation on the device (/root/reference/scripts/pyHICCUPS:143-166).
//
// From raw counts and the balancing weights it builds what the reference's worker hands to hiccups():
//   balanced diagonal d   = count * w[r] * w[r + d] where a count is stored (cooler's `balance=` transform,
//                           evaluated left to right), 0 elsewhere, NaN -> 0                  (:143, :153-157)
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
// (hp_hostpack.cpp) so that ~1 byte per band pixel crosses PCIe; this widens them back into the plain
// [num][pitch] int32 landing zone (zero tails) that k_relayout / k_prep_band read.  4 bins per thread.
struct PackedDiag {
    unsigned long long off;      // byte offset of the diagonal in the packed buffer (16-byte aligned)
    unsigned int esize;          // 1, 2 or 4
    unsigned int len;            // n - d
};

__global__ void __launch_bounds__(256) k_unpack_counts(const unsigned char* __restrict__ packed, const PackedDiag* __restrict__ tab,
                                                       int* __restrict__ raw_plain, int pitch) {
    const int d = blockIdx.y;
    const int rb = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (rb >= pitch) return;
    const PackedDiag t = tab[d];
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
    *reinterpret_cast<int4*>(raw_plain + (size_t)d * pitch + rb) = v;
}

// one CTA per diagonal d in [bal_first, num).  comp: scratch [num][pitch] doubles.  The summation tree has `depth`
// levels below the root and nslot = 2^(depth + 1) heap slots: in shared memory when it fits (use_smem), otherwise in
// the scratch arrays tree / tval / tlist, [num][nslot] each.
__global__ void __launch_bounds__(kPrepThreads) k_prep_band(const int* __restrict__ raw_plain, const double* __restrict__ w, int n, int num,
                                                            int pitch, int bal_first, double* __restrict__ bal, unsigned int* __restrict__ rownz,
                                                            double* __restrict__ ir, double* __restrict__ comp, int2* __restrict__ tree,
                                                            double* __restrict__ tval, int* __restrict__ tlist, int depth, int nslot,
                                                            int use_smem) {
    extern __shared__ __align__(16) unsigned char prep_smem[];     // use_smem: the summation tree (nodes, values, leaf list)
    __shared__ int sh_scan[kPrepThreads / 32];
    __shared__ int sh_base, sh_nleaf;
    const int d = bal_first + blockIdx.x;
    const int len = n - d;
    const int* src = raw_plain + (size_t)d * pitch;
    double* cp = comp + (size_t)blockIdx.x * pitch;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) sh_base = 0;
    __syncthreads();
    // ---- balanced values, NaN compaction (order preserved): 4 consecutive bins per thread and round ----------
    for (int r0 = 0; r0 < pitch; r0 += kPrepThreads * 4) {
        const int rb = r0 + threadIdx.x * 4;
        double v[4];
        bool keep[4];
        int nk = 0;
        int4 c4 = make_int4(0, 0, 0, 0);
        if (rb < pitch) c4 = *reinterpret_cast<const int4*>(src + rb);          // pitch is a multiple of 32
        const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = rb + k;
            const bool in = r < len;
            v[k] = 0.0;
            if (in) {                                   // weights loaded whether or not a count is stored: no load waits on another
                const double p = __dmul_rn(__dmul_rn((double)cc[k], w[r]), w[r + d]);
                if (cc[k] != 0) v[k] = p;
            }
            const bool isn = v[k] != v[k];
            keep[k] = in && !isn;
            nk += keep[k];
            if (r < pitch) {
                const double out = isn ? 0.0 : v[k];
                bal[qidx(d, r, pitch)] = out;
                if (out != 0.0) rownz[r] = 1u;
            }
        }
        int inc = nk;                                                            // inclusive scan of nk over the block
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) sh_scan[wid] = inc;
        __syncthreads();
        int pre = sh_base + inc - nk;
        for (int k = 0; k < wid; ++k) pre += sh_scan[k];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (keep[k]) cp[pre++] = v[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < kPrepThreads / 32; ++k) tot += sh_scan[k];
            sh_base += tot;
        }
        __syncthreads();
    }
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
            r = cp[L.x + j];
            for (int i = 8; i < L.y - (L.y % 8); i += 8) r = __dadd_rn(r, cp[L.x + i + j]);
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
