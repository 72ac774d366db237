// Exact re-evaluation of the records the fast score kernel (hp_score_fast.cuh) could not settle, and of the reported E
// of the survivors it classified.  The donut / lower-left sums are gathered from the balanced plane and added in the
// reference's fp64 order (the cell list of the sweep program, callers.py:147-198: steps in (w, p) order, rows a outer,
// columns b inner) -- the same list, walked in the same order, as the table-driven k_score.
#pragma once
#include "hp_kernels.cuh"
#include "hp_score_fast.cuh"

namespace hp {

constexpr int kExThreads = 128;
constexpr int kExRec = 16;          // records a warp evaluates side by side (k_fill_exact: thousands of records)
constexpr int kExRecFew = 4;        // ... in k_exact (a few hundred records: short critical path, every load of a round in flight)
constexpr int kExChunk = 128;       // cells gathered per round
constexpr int kExStride = kExChunk + 1;     // doubles between the cell buffers of two records: 16 records, 16 different bank pairs
__host__ __device__ constexpr size_t ex_warp_bytes(int nr) {      // cells, Y flags, (r, d, cells) per record
    return ((size_t)nr * kExStride * 8 + kExChunk + 3 * nr * 4 + 15) & ~(size_t)15;
}

// One warp, NR records (lane j < NR brings record j: row, diagonal, executed step; s < 0: no record).  Round by
// round the lanes gather kExChunk cells of the sweep's cell list for all the records (lane = cell: the cell's offsets are
// read once, the 16 loads of a lane are independent), then lane 2j walks the donut chain of record j and lane 2j + 1 its
// lower-left chain in list order -- 32 chains per fp64 instruction instead of one.  A cell outside a record's step range,
// the chromosome or the band, or outside the lower-left mask, enters the chain as +0.0, which leaves an fp64 sum of
// non-negative terms unchanged bit for bit.  Returns the sums of record `lane` (lanes < NR).
template <int NR>
__device__ __forceinline__ void exact_sums_batch(const Tables* __restrict__ tab, const double* __restrict__ bal, int n, int num, int pitch,
                                                 int bal_first, int myr, int myd, int mys, unsigned char* wbuf, int lane, double& SK,
                                                 double& SY) {
    const unsigned full = 0xffffffffu;
    double* const buf = reinterpret_cast<double*>(wbuf);
    unsigned char* const ys = wbuf + (size_t)NR * kExStride * 8;
    int* const sr = reinterpret_cast<int*>(ys + kExChunk);
    int* const sd = sr + NR;
    int* const sn = sd + NR;
    const int mynops = (lane < NR && mys >= 0) ? tab->prog.op_end[mys] : 0;
    if (lane < NR) { sr[lane] = myr; sd[lane] = myd; sn[lane] = mynops; }
    const int maxn = __reduce_max_sync(full, mynops);
    __syncwarp();
    const double* const chain = buf + (size_t)min(lane >> 1, NR - 1) * kExStride;
    const bool isy = (lane & 1) != 0;
    double acc = 0.0;
    for (int base = 0; base < maxn; base += kExChunk) {
        const int m = min(kExChunk, maxn - base);
        constexpr int CI = NR <= 4 ? kExChunk / 32 : 1;       // cell rounds whose loads are in flight together
        for (int i0 = 0; i0 < m; i0 += 32 * CI) {
            double v[CI][NR];
#pragma unroll
            for (int u = 0; u < CI; ++u) {
                const int i = i0 + 32 * u + lane, k = base + i;
                int a = 0, b = 0;
                if (i < m) { a = tab->opa[k]; b = tab->opb[k]; ys[i] = tab->opy[k]; }
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const int r = sr[j], d = sd[j];
                    const int rr = r + a, cc = r + d + b, dd = d + b - a;
                    v[u][j] = 0.0;
                    if (i < m && k < sn[j] && rr >= 0 && rr < n && cc >= 0 && cc < n && dd >= bal_first && dd < num)
                        v[u][j] = bal[qidx(dd, rr, pitch)];
                }
            }
#pragma unroll
            for (int u = 0; u < CI; ++u) {
                const int i = i0 + 32 * u + lane;
                if (i < m) {
#pragma unroll
                    for (int j = 0; j < NR; ++j) buf[(size_t)j * kExStride + i] = v[u][j];
                }
            }
        }
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            double v = chain[i];
            if (isy && !ys[i]) v = 0.0;
            acc = __dadd_rn(acc, v);
        }
        __syncwarp();
    }
    SK = __shfl_sync(full, acc, (2 * lane) & 31);
    SY = __shfl_sync(full, acc, (2 * lane + 1) & 31);
}

// records of the fast kernel's exact list -> the same per-pixel tail as the score kernels (emit_record)
__global__ void __launch_bounds__(kExThreads) k_exact(const __grid_constant__ ScoreArgs A, const double* __restrict__ bal,
                                                      const XRec* __restrict__ rec, const unsigned int* __restrict__ nrec_ptr,
                                                      unsigned int cap) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned nrec = *nrec_ptr;
    if (nrec > cap) nrec = cap;
    constexpr int NR = kExRecFew;
    const unsigned nbatch = (nrec + NR - 1) / NR;
    if (blockIdx.x >= nbatch) return;                   // batch b goes to warp b / gridDim.x of CTA b % gridDim.x
    const ScoreSmem sh = score_smem(smem, 0, 0, 0, 0, 0, 0);
    const int sh_bins = A.sh_pairs * 2 * kShI * kShK;
    unsigned char* bufs = smem + ((score_smem_bytes(0, 0, A.sh_pairs, 0, 0, 0, 0) + 127) & ~(size_t)127);
    if (threadIdx.x == 0) { *sh.cnt = 0; *sh.next = 0; }
    score_prologue<false>(A, sh, nullptr, 0, 0, 0, sh_bins);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbuf = bufs + (size_t)warp * ex_warp_bytes(NR);
    TailAcc tacc{};
    for (unsigned b = blockIdx.x + (unsigned)warp * gridDim.x; b < nbatch; b += gridDim.x * (kExThreads / 32)) {
        const unsigned k = b * NR + lane;
        const bool have = lane < NR && k < nrec;
        XRec x{0, 0, 0, 0};
        if (have) x = rec[k];
        const int d = x.ds & 0xffff, s = (x.ds >> 16) & 0xff;
        double sk, sy;
        exact_sums_batch<NR>(A.tab, bal, A.n, A.num, A.pitch, A.bal_first, x.r, d, have ? s : -1, wbuf, lane, sk, sy);
        emit_record<0, false>(A, sh, tacc, have, sk, sy, x.r, d, s, have ? A.step_pi[s] : 0, lane, 0, (unsigned)x.kind);
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
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbuf = smem + (size_t)warp * ex_warp_bytes(kExRec);
    unsigned ns = *A.nsurv_ptr;
    if (ns > A.cap) ns = A.cap;
    const unsigned nbatch = (ns + kExRec - 1) / kExRec;
    for (unsigned b = blockIdx.x + (unsigned)warp * gridDim.x; b < nbatch; b += gridDim.x * (kFillThreads / 32)) {
        const unsigned k = b * kExRec + lane;
        hp_survivor* sv = A.surv + k;
        int pr = 0, r = 0, d = 0, s = -1;
        if (lane < kExRec && k < ns) {
            pr = sv->pair;
            if (pr & kSurvNeedsE) { r = sv->r; d = sv->c - sv->r; s = (pr >> 8) & 0xff; }
        }
        if (__ballot_sync(0xffffffffu, s >= 0) == 0u) continue;      // no survivor of the fast list in this batch
        double SK, SY;
        exact_sums_batch<kExRec>(A.tab, A.bal, A.n, A.num, A.pitch, A.bal_first, r, d, s, wbuf, lane, SK, SY);
        if (s >= 0) {
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

// survivor selection over the fast kernel's candidate list: q <= sig for K or Y (callers.py:279-287).  q never rises with
// the observed count inside a lambda-chunk, so the test is "count >= the chunk's threshold" (kq, written by k_bh): the
// list (16 B per candidate, ~80 candidates per survivor) is streamed against a table in shared memory, and only the
// survivors touch the p / q tables.
constexpr int kFiltThreads = 256;
__global__ void __launch_bounds__(kFiltThreads) k_filter_fast(FilterArgs A, const FCand* __restrict__ fc, unsigned int nfc,
                                                              const int* __restrict__ kq, int npw) {
    __shared__ int s_kq[2 * HP_MAX_PW][kMaxChunk + 2], s_hw[kMaxChunk + 2], s_hoff[kMaxChunk + 2];
    __shared__ unsigned int s_nrej[2 * HP_MAX_PW];
    const Chunks& c_chunks = A.tab->chunks;
    for (int i = threadIdx.x; i < 2 * npw * (kMaxChunk + 2); i += kFiltThreads) {
        const int lf = i / (kMaxChunk + 2), ci = i % (kMaxChunk + 2);
        s_kq[lf][ci] = (ci >= 1 && ci <= A.numbin[lf]) ? kq[i] : 0x7fffffff;
        if (lf == 0) { s_hw[ci] = ci <= c_chunks.maxchunk ? c_chunks.hw[ci] : 1; s_hoff[ci] = ci <= c_chunks.maxchunk ? c_chunks.hoff[ci] : 0; }
    }
    if (threadIdx.x < 2 * HP_MAX_PW) s_nrej[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    for (unsigned base = blockIdx.x * kFiltThreads; base < nfc; base += gridDim.x * kFiltThreads) {      // warp-uniform trip count
        const unsigned idx = base + threadIdx.x;
        bool live = idx < nfc;
        FCand c{};
        if (live) *reinterpret_cast<int4*>(&c) = __ldcs(reinterpret_cast<const int4*>(&fc[idx]));
        if (c.r < 0) { live = false; c = FCand{}; }      // an unused slot of a warp's piece of the list
        const unsigned cflags = (c.info >> 16) & 0xffu;
        const int pair = (int)(c.info >> 24);
        const int ck = (int)(c.info & 0xffu), cy = (int)((c.info >> 8) & 0xffu);
        const int kbk = min(c.obs, s_hw[ck] - 1), kby = min(c.obs, s_hw[cy] - 1);
        const bool rk = live && (cflags & HP_SF_VALID_K) && kbk >= s_kq[pair * 2][ck];
        const bool ry = live && (cflags & HP_SF_VALID_Y) && kby >= s_kq[pair * 2 + 1][cy];
        if (rk) atomicAdd(&s_nrej[pair * 2], 1u);
        if (ry) atomicAdd(&s_nrej[pair * 2 + 1], 1u);
        const bool any = rk || ry;
        const unsigned many = __ballot_sync(0xffffffffu, any);
        if (many) {
            unsigned ob = 0;
            if (lane == (unsigned)__ffs(many) - 1u) ob = atomicAdd(&A.out_count[0], (unsigned)__popc(many));
            ob = __shfl_sync(0xffffffffu, ob, __ffs(many) - 1);
            if (any) {
                const int d = c.ds & 0xffff, s = (c.ds >> 16) & 0xff;
                hp_survivor sv;
                sv.r = c.r; sv.c = c.r + d; sv.pair = pair | (s << 8) | kSurvNeedsE;
                sv.flags = cflags | (rk ? HP_SF_REJECT_K : 0u) | (ry ? HP_SF_REJECT_Y : 0u);
                sv.obs = (double)c.obs;
                sv.e[0] = 0.0; sv.e[1] = 0.0;
                sv.p[0] = 1.0; sv.q[0] = 1.0; sv.p[1] = 1.0; sv.q[1] = 1.0;
                if (ck >= 1 && ck <= A.numbin[pair * 2]) {
                    sv.p[0] = A.ptab[s_hoff[ck] + kbk];
                    sv.q[0] = A.qtab[(size_t)(pair * 2) * c_chunks.total_bins + s_hoff[ck] + kbk];
                }
                if (cy >= 1 && cy <= A.numbin[pair * 2 + 1]) {
                    sv.p[1] = A.ptab[s_hoff[cy] + kby];
                    sv.q[1] = A.qtab[(size_t)(pair * 2 + 1) * c_chunks.total_bins + s_hoff[cy] + kby];
                }
                sv.ice = A.bal[qidx(d, c.r, A.pitch)];
                const unsigned g = ob + __popc(many & ((1u << lane) - 1u));
                if (g < A.out_cap) A.out[g] = sv; else atomicAdd(&A.out_count[1], 1u);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * npw && s_nrej[threadIdx.x]) atomicAdd(&A.nreject[threadIdx.x], (unsigned long long)s_nrej[threadIdx.x]);
}

}  // namespace hp
