// Exact re-evaluation of the records the fast score kernel (hp_score_fast.cuh) could not settle, and of the reported E
// of the survivors it classified.  The donut / lower-left sums are gathered from the balanced plane and added in the
// reference's fp64 order (the cell list of the sweep program, callers.py:147-198: steps in (w, p) order, rows a outer,
// columns b inner) -- the same list, walked in the same order, as the table-driven k_score.
#pragma once
#include "hp_kernels.cuh"
#include "hp_score_fast.cuh"

namespace hp {

constexpr int kExThreads = 256;
constexpr int kExChunk = 256;       // cells gathered per round and warp

// One warp, one record: the lanes gather the cells of steps 0..s (independent loads), lane 0 adds the donut sum and lane 1
// the lower-left sum in list order.  A cell that is not part of the lower-left mask enters that chain as +0.0, which
// leaves an fp64 sum of non-negative terms unchanged bit for bit.
__device__ __forceinline__ void exact_sums_warp(const Tables* __restrict__ tab, const double* __restrict__ bal, int n, int num, int pitch,
                                                int bal_first, int r, int d, int s, double* buf, double* bufy, int lane, double& SK,
                                                double& SY) {
    const int nops = tab->prog.op_end[s];
    double sk = 0.0, sy = 0.0;
    for (int base = 0; base < nops; base += kExChunk) {
        const int m = min(kExChunk, nops - base);
        for (int i0 = lane; i0 < m; i0 += 128) {          // four independent loads in flight per lane
            double v[4];
            bool y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 32 * u;
                v[u] = 0.0; y[u] = false;
                if (i < m) {
                    const int k = base + i;
                    const int a = tab->opa[k], b = tab->opb[k];
                    const int rr = r + a, cc = r + d + b, dd = d + b - a;
                    y[u] = tab->opy[k] != 0;
                    if (rr >= 0 && rr < n && cc >= 0 && cc < n && dd >= bal_first && dd < num) v[u] = bal[qidx(dd, rr, pitch)];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 32 * u;
                if (i < m) { buf[i] = v[u]; bufy[i] = y[u] ? v[u] : 0.0; }
            }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll 4
            for (int i = 0; i < m; ++i) sk = __dadd_rn(sk, buf[i]);
        } else if (lane == 1) {
#pragma unroll 4
            for (int i = 0; i < m; ++i) sy = __dadd_rn(sy, bufy[i]);
        }
        __syncwarp();
    }
    SK = __shfl_sync(0xffffffffu, sk, 0);
    SY = __shfl_sync(0xffffffffu, sy, 1);
}

// records of the fast kernel's exact list -> the same per-pixel tail as the score kernels (emit_record)
__global__ void __launch_bounds__(kExThreads) k_exact(const __grid_constant__ ScoreArgs A, const double* __restrict__ bal,
                                                      const XRec* __restrict__ rec, const unsigned int* __restrict__ nrec_ptr,
                                                      unsigned int cap) {
    extern __shared__ __align__(128) unsigned char smem[];
    {
        unsigned nr = *nrec_ptr;
        if (nr > cap) nr = cap;
        const unsigned nw = gridDim.x * (kExThreads / 32), pw = (nr + nw - 1) / nw;
        if ((unsigned long long)blockIdx.x * (kExThreads / 32) * pw >= nr) return;      // no record for this CTA
    }
    const ScoreSmem sh = score_smem(smem, 0, 0, 0, 0, 0, 0);
    const int sh_bins = A.sh_pairs * 2 * kShI * kShK;
    double* bufs = reinterpret_cast<double*>(smem + ((score_smem_bytes(0, 0, A.sh_pairs, 0, 0, 0, 0) + 127) & ~(size_t)127));
    if (threadIdx.x == 0) { *sh.cnt = 0; *sh.next = 0; }
    score_prologue<false>(A, sh, nullptr, 0, 0, 0, sh_bins);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf = bufs + (size_t)warp * 2 * kExChunk;
    double* bufy = buf + kExChunk;
    unsigned nrec = *nrec_ptr;
    if (nrec > cap) nrec = cap;
    const unsigned nwarps = gridDim.x * (kExThreads / 32), gw = blockIdx.x * (kExThreads / 32) + warp;
    const unsigned per = (nrec + nwarps - 1) / nwarps;
    const unsigned begin = min(nrec, gw * per), endr = min(nrec, begin + per);
    TailAcc tacc{};
    for (unsigned b0 = begin; b0 < endr; b0 += 32) {
        const int nb = (int)min(32u, endr - b0);
        double mySK = 0.0, mySY = 0.0;
        int myr = 0, myd = 0, mys = 0;
        unsigned mykind = 0;
        for (int j = 0; j < nb; ++j) {
            const XRec x = rec[b0 + j];
            const int d = x.ds & 0xffff, s = (x.ds >> 16) & 0xff;
            double sk, sy;
            exact_sums_warp(A.tab, bal, A.n, A.num, A.pitch, A.bal_first, x.r, d, s, buf, bufy, lane, sk, sy);
            if (lane == j) { mySK = sk; mySY = sy; myr = x.r; myd = d; mys = s; mykind = (unsigned)x.kind; }
        }
        emit_record<0, false>(A, sh, tacc, lane < nb, mySK, mySY, myr, myd, mys, 0, lane, 0, mykind);
    }
    score_epilogue(A, sh, sh_bins);
}

// The fast kernel's candidates carry no E.  After the FDR step the survivors among them (flag bit 30 of `pair`, step in
// bits 8..15) get E, the validity flags and cEM != 0 from the exact sums, exactly as emit_record computes them.
struct FillArgs {
    const Tables* tab;
    const double* bal;
    const double* ir;
    const double* b1;
    const double* b2;
    const double* betab;
    hp_survivor* surv;
    const unsigned int* nsurv_ptr;
    unsigned int cap;
    int n, num, pitch, bal_first, F, nexec;
};
constexpr int kSurvNeedsE = 1 << 30;

constexpr int kFillThreads = 128;
__global__ void __launch_bounds__(kFillThreads) k_fill_exact(const FillArgs A) {
    __shared__ double bufs[(kFillThreads / 32) * 2 * kExChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf = bufs + (size_t)warp * 2 * kExChunk;
    double* bufy = buf + kExChunk;
    unsigned ns = *A.nsurv_ptr;
    if (ns > A.cap) ns = A.cap;
    const unsigned nwarps = gridDim.x * (kFillThreads / 32);
    for (unsigned k = blockIdx.x * (kFillThreads / 32) + warp; k < ns; k += nwarps) {
        hp_survivor* sv = A.surv + k;
        const int pr = sv->pair;
        if (!(pr & kSurvNeedsE)) continue;               // warp-uniform: every lane reads the same record
        const int r = sv->r, d = sv->c - sv->r, s = (pr >> 8) & 0xff;
        double SK, SY;
        exact_sums_warp(A.tab, A.bal, A.n, A.num, A.pitch, A.bal_first, r, d, s, buf, bufy, lane, SK, SY);
        if (lane == 0) {
            double be[2];
            record_be(BeArgs{A.tab, A.ir, A.betab, A.n, A.num, A.F, A.nexec, A.bal_first}, r, d, s, be);
            const RecVal V = record_values(SK, SY, be, A.ir[d], A.b1[r], A.b2[r + d], true);
            unsigned fl = sv->flags & ~(HP_SF_VALID_K | HP_SF_VALID_Y | HP_SF_CEMY_NONZERO);
            if (V.valid[0]) fl |= HP_SF_VALID_K;
            if (V.valid[1]) fl |= HP_SF_VALID_Y;
            if (V.cnz[1]) fl |= HP_SF_CEMY_NONZERO;
            sv->flags = fl;
            sv->e[0] = V.valid[0] ? V.E[0] : 0.0;
            sv->e[1] = V.valid[1] ? V.E[1] : 0.0;
            sv->pair = pr & 0xff;
        }
        __syncwarp();
    }
}

// survivor selection over the fast kernel's candidate list: q <= sig for K or Y (callers.py:279-287)
__global__ void k_filter_fast(FilterArgs A, const FCand* __restrict__ fc, unsigned int nfc) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = idx < nfc;
    const unsigned lane = threadIdx.x & 31u;
    const Chunks& c_chunks = A.tab->chunks;
    FCand c{};
    if (live) *reinterpret_cast<int4*>(&c) = *reinterpret_cast<const int4*>(&fc[idx]);
    if (c.r < 0) { live = false; c = FCand{}; }           // an unused slot of a warp's piece of the list
    const int d = c.ds & 0xffff, s = (c.ds >> 16) & 0xff;
    const unsigned cflags = (c.info >> 16) & 0xffu;
    const int pair = (int)(c.info >> 24);
    hp_survivor sv;
    sv.r = c.r; sv.c = c.r + d; sv.pair = pair | (s << 8) | kSurvNeedsE; sv.flags = cflags;
    sv.obs = (double)c.obs;
    sv.e[0] = 0.0; sv.e[1] = 0.0;
    bool rej[2] = {false, false};
#pragma unroll
    for (int fl = 0; fl < 2; ++fl) {
        const int lf = pair * 2 + fl;
        const int ci = (int)((c.info >> (8 * fl)) & 0xffu);
        double p = 1.0, q = 1.0;
        if (live && ci >= 1 && ci <= A.numbin[lf]) {
            const int w = c_chunks.hw[ci];
            const int kb = c.obs < w - 1 ? c.obs : w - 1;
            p = A.ptab[c_chunks.hoff[ci] + kb];
            q = A.qtab[(size_t)lf * c_chunks.total_bins + c_chunks.hoff[ci] + kb];
        }
        sv.p[fl] = p; sv.q[fl] = q;
        const bool valid = (cflags & (fl ? HP_SF_VALID_Y : HP_SF_VALID_K)) != 0;
        rej[fl] = live && valid && q <= A.sig;
    }
    if (rej[0]) sv.flags |= HP_SF_REJECT_K;
    if (rej[1]) sv.flags |= HP_SF_REJECT_Y;
#pragma unroll
    for (int fl = 0; fl < 2; ++fl) {
        const unsigned peers = __match_any_sync(0xffffffffu, rej[fl] ? pair : -1);
        if (rej[fl] && lane == (unsigned)__ffs(peers) - 1u) atomicAdd(&A.nreject[pair * 2 + fl], (unsigned long long)__popc(peers));
    }
    const bool any = rej[0] || rej[1];
    const unsigned many = __ballot_sync(0xffffffffu, any);
    if (many) {
        unsigned base = 0;
        if (lane == (unsigned)__ffs(many) - 1u) base = atomicAdd(&A.out_count[0], (unsigned)__popc(many));
        base = __shfl_sync(0xffffffffu, base, __ffs(many) - 1);
        if (any) {
            sv.ice = A.bal[qidx(d, c.r, A.pitch)];
            const unsigned g = base + __popc(many & ((1u << lane) - 1u));
            if (g < A.out_cap) A.out[g] = sv; else atomicAdd(&A.out_count[1], 1u);
        }
    }
}

}  // namespace hp
