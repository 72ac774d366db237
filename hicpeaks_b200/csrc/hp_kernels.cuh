// sm_100a kernels of the HiCCUPS scoring path.  Band layout in HBM: diagonal-major planes
// X[d][r] = matrix element (r, r + d), row pitch `pitch` elements (multiple of 32), zero beyond the
// chromosome end; shift(X, a, b)[r, c] = X[r + a, c + b] is plane element [d + b - a][r + a].
#pragma once
#include "hp_device.cuh"
#include "hp_poisson.cuh"

namespace hp {

// ============================================================================================
// K1  level kernel (generic) -- callers.py:197-198 (Reads accumulation) + :203-206 (Reads >= min_local_reads)
// For every non-zero band pixel: the first sweep step s* after which the raw lower-left sum reaches
// the threshold.  Integer work on the raw plane only; one TMA tile (+ lower-left halo) per CTA.
// The specialised version for compiled-in sweep programs is k_levels_spec (hp_score_spec.cuh).
// ============================================================================================
struct LevelArgs {
    const Tables* tab;
    unsigned char* lvl;            // quad-interleaved [num][pitch]
    unsigned long long* hist;      // [nsteps + 1]   (index nsteps = never)
    int n, pitch, dlo, dhi, F, BD, TD, NQ;
};

__device__ __forceinline__ void level_prologue(int* tile, uint64_t* bar, const CUtensorMap* tm, int tile_bytes, int q0, int plane0) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)tile_bytes);
        tma_load_3d(tile, tm, q0, 0, plane0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
}

__global__ void __launch_bounds__(kThreads) k_levels(const __grid_constant__ CUtensorMap tm_raw, LevelArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Tables& T = *A.tab;
    int* tile = reinterpret_cast<int*>(smem);                       // [BD][4][NQ], rows r0.., planes d0 - 2F..
    const int tile_bytes = A.BD * 4 * A.NQ * 4;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + tile_bytes);
    unsigned int* sh_hist = reinterpret_cast<unsigned int*>(smem + tile_bytes + 16);

    const int r0 = blockIdx.x * kTR;
    const int d0 = A.dlo + blockIdx.y * A.TD;
    const int nsteps = T.prog.nsteps;
    for (int i = threadIdx.x; i <= nsteps; i += kThreads) sh_hist[i] = 0;
    level_prologue(tile, bar, &tm_raw, tile_bytes, r0 / 4, d0 - 2 * A.F);

    const int rl = threadIdx.x & (kTR - 1);
    const int r = r0 + rl;
    const long long thr = T.prog.thr;
    for (int dl = threadIdx.x >> 7; dl < A.TD; dl += kThreads / kTR) {
        const int d = d0 + dl;
        if (d > A.dhi || r >= A.n) continue;
        unsigned char out = kLvlNone;
        if (r + d < A.n && tile[((dl + 2 * A.F) * 4 + (rl & 3)) * A.NQ + (rl >> 2)] != 0) {
            long long R = 0;
            int opi = 0;
            out = kLvlNever;
            for (int s = 0; s < nsteps; ++s) {
                const int e = T.prog.rop_end[s];
                for (; opi < e; ++opi) {
                    const int k = T.ropi[opi];
                    const int a = T.opa[k], b = T.opb[k], rr = rl + a;
                    R += tile[((dl + 2 * A.F + b - a) * 4 + (rr & 3)) * A.NQ + (rr >> 2)];
                }
                if (R >= thr) { out = (unsigned char)s; break; }
            }
            const int slot = (out == kLvlNever) ? nsteps : out;
            const unsigned m = __match_any_sync(__activemask(), slot);
            if ((threadIdx.x & 31) == __ffs(m) - 1) smem_red_add(&sh_hist[slot], __popc(m));
        }
        A.lvl[qidx(d, r, A.pitch)] = out;
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= nsteps; i += kThreads)
        if (sh_hist[i]) atomicAdd(&A.hist[i], (unsigned long long)sh_hist[i]);
}

// ============================================================================================
// expected-sum table for interior pixels: bE[fl][s][d] = ordered sum of IR over the offsets added
// up to and including step s (callers.py:178,182,189-191).  Away from the chromosome ends the sum
// only depends on the diagonal, so it never touches the band.
// ============================================================================================
// Layout betab[z][fl][s][d]: z = 0 interior pixels; z = 1 + r for the rows r < F next to the start of the
// chromosome (cells with row < 0 or column < 0 drop out); z = 1 + F + e for the columns c = n - 1 - e, e < F,
// next to its end (cells with row >= n or column >= n drop out).  Pixels near both ends take edge_be().
// ffac (optional): the same table as fp32 factors IR[d] / bE for the re-associated kernel (hp_score_fast.cuh);
// ffs: its interior part again, per pair, as [strip][fl][code][64] blocks (one bulk copy per tile; code = width of the
// step - ww of its pair, nc codes per pair)
struct FfsLayout {
    float* base;                       // nullptr: none
    int off[HP_MAX_PW];                // first float of the pair's blocks
    int nc[HP_MAX_PW];                 // codes of the pair among the executed steps
};
constexpr int kBetabThreads = 64;
__global__ void __launch_bounds__(kBetabThreads) k_betab(const Tables* __restrict__ tab, const double* __restrict__ ir, double* __restrict__ betab,
                                                         int num, int bal_first, int nsteps_exec, int F, float* __restrict__ ffac,
                                                         const FfsLayout ffs) {
    // One thread per (diagonal, edge variant) walks the cell list ONCE and leaves its running sums at every step boundary
    // (the accumulators are never reset between steps, callers.py:132-198).  The cells are staged in shared memory,
    // decoded; the two fp64 chains stay sequential, everything around them is independent and four cells wide.
    __shared__ int pk[kMaxOps];                         // (b - a) + 64 | a + 64 << 8 | b + 64 << 16 | y << 24
    const int nops = tab->prog.op_end[nsteps_exec - 1];
    for (int i = threadIdx.x; i < nops; i += kBetabThreads) {
        const int a = tab->opa[i], b = tab->opb[i];
        pk[i] = (b - a + 64) | ((a + 64) << 8) | ((b + 64) << 16) | ((tab->opy[i] ? 1 : 0) << 24);
    }
    __syncthreads();
    const int d = blockIdx.x * kBetabThreads + threadIdx.x;
    const int z = blockIdx.y;
    if (d >= num) return;
    int amin = -128, bmin = -128, amax = 127, bmax = 127;
    if (z >= 1 && z <= F) { const int r = z - 1; amin = -r; bmin = -(r + d); }
    if (z > F) { const int e = z - F - 1; amax = e + d; bmax = e; }
    const double ird = ir[d];
    double ek = 0.0, ey = 0.0;
    int i = 0;
    for (int s = 0; s < nsteps_exec; ++s) {
        const int e = tab->prog.op_end[s];
        for (; i < e; i += 4) {
            double v[4];
            bool on[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int w = (i + u < e) ? pk[i + u] : 0;
                const int dd = d + (w & 0xff) - 64, a = ((w >> 8) & 0xff) - 64, b = ((w >> 16) & 0xff) - 64;
                on[u] = (i + u < e) && dd >= bal_first && dd < num && a >= amin && a <= amax && b >= bmin && b <= bmax;
                y[u] = (w >> 24) != 0;
                v[u] = on[u] ? ir[dd] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (on[u]) {
                    ek = __dadd_rn(ek, v[u]);
                    if (y[u]) ey = __dadd_rn(ey, v[u]);
                }
            }
        }
        i = e;
        betab[((size_t)(z * 2 + 0) * nsteps_exec + s) * num + d] = ek;
        betab[((size_t)(z * 2 + 1) * nsteps_exec + s) * num + d] = ey;
        if (ffac) {
            const float fk = fast_factor(ird, ek), fy = fast_factor(ird, ey);
            ffac[((size_t)(z * 2 + 0) * nsteps_exec + s) * num + d] = fk;
            ffac[((size_t)(z * 2 + 1) * nsteps_exec + s) * num + d] = fy;
            if (ffs.base && z == 0 && d >= bal_first) {     // the interior factors once more, one block of 64 diagonals per strip
                const int k = d - bal_first, pi = tab->prog.step_pi[s], nc = ffs.nc[pi];
                const int code = tab->prog.step_w[s] - tab->prog.ww[pi];
                float* o = ffs.base + ffs.off[pi] + ((size_t)(k >> 6) * 2 * nc + code) * 64 + (k & 63);
                o[0] = fk;
                o[(size_t)nc * 64] = fy;
            }
        }
    }
}

// same sum for a pixel next to a chromosome end (some offsets fall outside [0, n)^2)
__device__ __noinline__ void edge_be(const Tables* __restrict__ tab, const double* __restrict__ ir, int r, int d, int n, int num,
                                     int bal_first, int s, double& ek, double& ey) {
    ek = 0.0; ey = 0.0;
    const int e = tab->prog.op_end[s];
    const int c = r + d;
    for (int i = 0; i < e; ++i) {
        const int a = tab->opa[i], b = tab->opb[i];
        const int dd = d + b - a, rr = r + a, cc = c + b;
        if (rr >= 0 && rr < n && cc >= 0 && cc < n && dd >= bal_first && dd < num) {
            const double v = ir[dd];
            ek = __dadd_rn(ek, v);
            if (tab->opy[i]) ey = __dadd_rn(ey, v);
        }
    }
}

// ============================================================================================
// K2  score kernels -- callers.py:132-198 (donut / lower-left sums in the reference's fp64 order),
// :212-213 (snapshot at the resolving step), :244-256 (E, validity), :25-41 (lambda-chunk id) and
// the (chunk, observed) histogram that replaces the per-chunk sort of multipletests (:265-275).
// One TMA tile of the quad-interleaved balanced plane with halo per CTA.  Two kernels share the
// per-pixel tail `emit_pixel`: k_score (any sweep program, table-driven) and k_score_spec
// (hp_score_spec.cuh: compile-time unrolled sweep programs, 4x4 register blocks).
// ============================================================================================
struct ScoreArgs {
    const Tables* tab;
    const int* raw;                    // quad-interleaved
    const unsigned char* lvl;          // quad-interleaved
    const double* ir;
    const double* b1;
    const double* b2;
    const double* betab;               // [1 + 2F][2][nsteps_exec][num], see k_betab
    unsigned int* hist;                // [npw*2][total_bins]
    unsigned long long* emax_bits;     // [npw*2]
    unsigned long long* nvalid;        // [npw*2]
    Cand* cand;
    unsigned int* cand_count;          // [0] records, [1] dropped (capacity), [2] chunk overflow
    unsigned int cand_cap;
    double* dump;                      // optional [npw*2][3][num*pitch]
    long long plane;
    int n, num, pitch, dlo, dhi, F, BD, TD, bal_first, sh_pairs;
    int HR, NQ;                        // row halo (multiple of 8) and row quads of the tile: NQ = (kTR + 2 HR) / 4
    // hot scalars and step tables of the program, in the kernel parameter (constant) bank
    int nexec, npw, dspan, maxchunk, total_bins;
    int bhfdr;                         // BH-FDR caller (callers.py:364-553): donut only, per-pixel Poisson rate, no histograms
    int ww[HP_MAX_PW];
    unsigned char step_pi[HP_MAX_STEPS], step_lo[HP_MAX_STEPS];
    unsigned char last_need[HP_MAX_WW + 1][HP_MAX_STEPS + 2];
};

struct ScoreSmem {                     // carve-up of the dynamic shared memory of a score CTA
    double* tile;                      // [BD][4][NQ]
    uint64_t* bar;
    unsigned int* cnt;                 // staged candidate count
    unsigned int* base;
    unsigned int* next;                // next unclaimed column block of the CTA (k_score_spec)
    unsigned long long* emax;          // [16]
    unsigned int* nval;                // [16]
    double* rv;                        // [kChunkTab] chunk edges (copy of c_chunks.rv: per-lane indexing)
    int4* cinfo;                       // [kChunkTab] {hoff, hw, kcand, 0}
    Cand* stage;                       // [kStage]
    double2* qsum;                     // [warps][kQCap] resolved (bS_K, bS_Y) waiting for the tail (k_score_spec only)
    int2* qmeta;                       // [warps][kQCap] {r, d << 16 | step << 8 | pair}
    int* qobs;                         // [warps][kQCap] raw count of the pixel
    // per-CTA copies of the small tables the tail reads (k_score_spec only): interior bE, IR, biases
    double* tb_be;                     // [2][nexec][tb_nd], diagonals tb_d0 .. tb_d0 + tb_nd - 1
    double* tb_ir;                     // [tb_nd]
    double* tb_b1;                     // [kTR] rows r0 ..
    double* tb_b2;                     // [kTR + tb_nd] columns r0 + tb_d0 ..
    int tb_d0, tb_nd, tb_r0;
    unsigned int* hist;                // [sh_pairs*2][kShI][kShK]
};
__host__ __device__ __forceinline__ size_t score_tab_bytes(int nexec, int nd) {
    return ((size_t)2 * nexec * nd + nd + kTR + kTR + nd) * 8;
}
constexpr int kChunkTab = kMaxChunk + 4;   // shared-memory copies of the chunk tables, +inf padded
constexpr int kChunkTabBytes = kChunkTab * 24;
// qwarps: warps that own a tail queue (0: none), qcap: records per queue
// tab_nexec / tab_nd: size the per-CTA table copies (0: none)
__host__ __device__ __forceinline__ size_t score_smem_bytes(int BD, int NQ, int sh_pairs, int qwarps, int qcap, int tab_nexec, int tab_nd) {
    return (size_t)BD * 4 * NQ * 8 + 48 + 128 + 64 + kChunkTabBytes + (size_t)kStage * sizeof(Cand) + (size_t)qwarps * qcap * 28 +
           (tab_nd ? score_tab_bytes(tab_nexec, tab_nd) : 0) + (size_t)sh_pairs * 2 * kShI * kShK * 4;
}
__device__ __forceinline__ ScoreSmem score_smem(unsigned char* smem, int BD, int NQ, int qwarps, int qcap, int tab_nexec, int tab_nd) {
    ScoreSmem S;
    unsigned char* p = smem;
    S.tile = reinterpret_cast<double*>(p); p += (size_t)BD * 4 * NQ * 8;
    S.bar = reinterpret_cast<uint64_t*>(p);                       // kTileParts barriers
    S.cnt = reinterpret_cast<unsigned int*>(p + 32);
    S.base = reinterpret_cast<unsigned int*>(p + 36);
    S.next = reinterpret_cast<unsigned int*>(p + 40); p += 48;
    S.emax = reinterpret_cast<unsigned long long*>(p); p += 128;
    S.nval = reinterpret_cast<unsigned int*>(p); p += 64;
    S.cinfo = reinterpret_cast<int4*>(p); p += kChunkTab * 16;
    S.rv = reinterpret_cast<double*>(p); p += kChunkTab * 8;
    S.stage = reinterpret_cast<Cand*>(p); p += (size_t)kStage * sizeof(Cand);
    S.qsum = reinterpret_cast<double2*>(p);
    S.qmeta = reinterpret_cast<int2*>(p + (size_t)qwarps * qcap * 16);
    S.qobs = reinterpret_cast<int*>(p + (size_t)qwarps * qcap * 24);
    p += (size_t)qwarps * qcap * 28;
    p += (16 - ((size_t)(p - smem) & 15)) & 15;
    S.tb_be = reinterpret_cast<double*>(p);
    S.tb_ir = S.tb_be + (size_t)2 * tab_nexec * tab_nd;
    S.tb_b1 = S.tb_ir + tab_nd;
    S.tb_b2 = S.tb_b1 + kTR;
    S.tb_d0 = 0; S.tb_nd = tab_nd; S.tb_r0 = 0;
    if (tab_nd) p += score_tab_bytes(tab_nexec, tab_nd);
    S.hist = reinterpret_cast<unsigned int*>(p);
    return S;
}

// lambda-chunk of E > 0: smallest i >= 1 with E < rv[i]; member iff rv[i-1] < E (strict, callers.py:38).
// hp_ctx_create checks that every third edge is an exact power of two (rv[3e+1] = 2^e, as numpy and the C
// library both give), so for 2^e <= E < 2^(e+1) the chunk is 3e+2, 3e+3 or 3e+4: two compares, no search.
// rv[] is padded with +inf beyond maxchunk.  Returns maxchunk + 1 when E is beyond the last edge.
__device__ __forceinline__ int find_chunk(const double* __restrict__ rv, int mc, double E, bool& member) {
    int i = 1;
    double lower = 0.0;
    if (E >= 1.0) {
        const int e = (int)((__double2hiint(E) >> 20) & 0x7ff) - 1023;
        int i0 = 3 * e + 2;
        if (i0 > mc + 1) i0 = mc + 1;
        const double e0 = rv[i0], e1 = rv[i0 + 1];
        lower = rv[i0 - 1];
        i = i0;
        if (E >= e0) { lower = e0; i = i0 + 1; }
        if (E >= e1) { lower = e1; i = i0 + 2; }
    }
    member = (i <= mc) && (E > lower);
    return i;
}

struct TailAcc {                       // per-thread running totals of a single-pair kernel (merged at the end)
    unsigned long long emax[2];
    unsigned int nval[2];
};

// bE of a record from the tables in global memory (interior / next to one chromosome end) or from the cell list
// (next to both ends) -- callers.py:178,182,189-191
struct BeArgs {
    const Tables* tab;
    const double* ir;
    const double* betab;
    int n, num, F, nexec, bal_first;
};
__device__ __forceinline__ void record_be(const BeArgs& A, int r, int d, int s, double (&be)[2]) {
    const bool top = r < A.F, end = r + d >= A.n - A.F;
    if (top && end) {                          // chromosome shorter than the band + two windows: walk the cell list
        edge_be(A.tab, A.ir, r, d, A.n, A.num, A.bal_first, s, be[0], be[1]);
    } else {
        const int z = top ? 1 + r : end ? 1 + A.F + (A.n - 1 - r - d) : 0;
        const double* bt = A.betab + ((size_t)(z * 2) * A.nexec + s) * A.num + d;
        be[0] = bt[0];
        be[1] = bt[(size_t)A.nexec * A.num];
    }
}

// callers.py:244-253: E = ((IR[d] * (bS / bE)) * B1[r]) * B2[c] for the donut (0) and the lower-left (1) background,
// cnz = the reference's cEM entry is stored (non-zero), valid = E > 0.  The two halves are independent and evaluated
// side by side so that the two division chains overlap.  Shared by the score kernels, k_exact and k_fill_exact: one
// sequence of operations, one result.
struct RecVal {
    double E[2];
    bool cnz[2], valid[2];
};
__device__ __forceinline__ RecVal record_values(double SK, double SY, const double (&be)[2], double ird, double bb1, double bb2, bool two) {
    RecVal R;
    const double den0 = be[0] != 0.0 ? be[0] : 1.0, den1 = be[1] != 0.0 ? be[1] : 1.0;
    const double ratio0 = __ddiv_rn(SK, den0), ratio1 = __ddiv_rn(SY, den1);
    const double cem0 = __dmul_rn(ird, ratio0), cem1 = __dmul_rn(ird, ratio1);
    R.E[0] = __dmul_rn(__dmul_rn(cem0, bb1), bb2);
    R.E[1] = __dmul_rn(__dmul_rn(cem1, bb1), bb2);
    R.cnz[0] = (be[0] != 0.0) && (ratio0 != 0.0) && (cem0 != 0.0);
    R.cnz[1] = two && (be[1] != 0.0) && (ratio1 != 0.0) && (cem1 != 0.0);
    R.valid[0] = R.cnz[0] && (R.E[0] > 0.0);
    R.valid[1] = R.cnz[1] && (R.E[1] > 0.0);
    return R;
}

// Per-pixel tail (callers.py:244-256 + chunk id + histograms), called by every lane of a converged warp.
// `act`: this lane holds a pixel (r, r + d) that resolves pair `pi` at executed step `s` with donut /
// lower-left sums SK, SY; s and pi may differ between lanes.  NPW == 1: single pair, totals kept in `acc`.
// SM: the small tables come from the CTA's shared-memory copies and the raw count from `obs_in`.
// kind (k_exact only): 0 = account the record; bit 0 / 1 = the record was accounted by the fast kernel, only
// E.max() of K / Y takes its exact value.
template <int NPW, bool SM>
__device__ __forceinline__ void emit_record(const ScoreArgs& A, const ScoreSmem& sh, TailAcc& acc, bool act, double SK, double SY,
                                            int r, int d, int s, int pi, int lane, int obs_in, unsigned kind = 0u) {
    const int nexec = A.nexec;
    const int mc = A.maxchunk;
    if (NPW == 1) pi = 0;
    bool cand = false;
    unsigned flags = 0, chk[2] = {0, 0};
    double Ev[2] = {0.0, 0.0};
    int obs = 0;
    const bool acct = kind == 0u;
    if (act) {
        obs = SM ? obs_in : A.raw[qidx(d, r, A.pitch)];
        double be[2];
        const bool top = r < A.F, end = r + d >= A.n - A.F;
        if (SM && !top && !end) {
            const double* bt = sh.tb_be + (size_t)s * sh.tb_nd + (d - sh.tb_d0);
            be[0] = bt[0];
            be[1] = bt[(size_t)nexec * sh.tb_nd];
        } else {
            record_be(BeArgs{A.tab, A.ir, A.betab, A.n, A.num, A.F, A.nexec, A.bal_first}, r, d, s, be);
        }
        const double ird = SM ? sh.tb_ir[d - sh.tb_d0] : A.ir[d];
        const double bb1 = SM ? sh.tb_b1[r - sh.tb_r0] : A.b1[r];
        const double bb2 = SM ? sh.tb_b2[r + d - sh.tb_r0 - sh.tb_d0] : A.b2[r + d];
        // The donut (K) and lower-left (Y) halves of a record are independent until they touch the histograms.  They
        // are evaluated side by side in straight-line code -- both divisions, then both chunk searches -- so that the
        // two dependent chains (division: ~60 cycles) overlap instead of running one after the other behind branches;
        // with four warps per scheduler the length of this tail, not its instruction count, is what keeps the fp64
        // pipe idle.  Results of a half that turns out invalid are discarded, exactly as the branches did.
        const bool two = !A.bhfdr;                  // the BH-FDR caller has no lower-left background (callers.py:440-540)
        const RecVal V = record_values(SK, SY, be, ird, bb1, bb2, two);
        bool memf[2];
        int cif[2];
        int4 inff[2];
        cif[0] = find_chunk(sh.rv, mc, V.valid[0] ? V.E[0] : 0.5, memf[0]);
        cif[1] = find_chunk(sh.rv, mc, V.valid[1] ? V.E[1] : 0.5, memf[1]);
        inff[0] = sh.cinfo[cif[0]];                 // cinfo / rv are padded beyond maxchunk (score_prologue)
        inff[1] = sh.cinfo[cif[1]];
        if (V.cnz[1] && acct) flags |= HP_SF_CEMY_NONZERO;
#pragma unroll
        for (int fl = 0; fl < 2; ++fl) {
            const double E = V.E[fl];
            const bool valid = V.valid[fl];
            if (A.dump && (fl == 0 || two)) {
                double* dp = A.dump + (size_t)((pi * 2 + fl) * 3) * A.plane + (size_t)d * A.pitch + r;
                dp[0] = fl ? SY : SK;
                dp[A.plane] = be[fl];
                dp[2 * A.plane] = valid ? E : 0.0;
            }
            if (valid) {
                const unsigned long long eb = (unsigned long long)__double_as_longlong(E);
                if (acct || ((kind >> fl) & 1u)) {
                    if (NPW == 1) acc.emax[fl] = eb > acc.emax[fl] ? eb : acc.emax[fl];
                    else if (eb > sh.emax[pi * 2 + fl]) smem_red_max(&sh.emax[pi * 2 + fl], eb);
                }
                if (!acct) continue;
                flags |= (fl ? HP_SF_VALID_Y : HP_SF_VALID_K);
                Ev[fl] = E;
                if (NPW == 1) ++acc.nval[fl];
                const int ci = cif[fl];
                if (ci > mc) atomicAdd(&A.cand_count[2], 1u);
                if (A.bhfdr) {
                    // p = 1 - pdtr(O, E) grows with E: with E in [rv[ci-1], rv[ci]) it can only pass sig if the tail at
                    // the chunk's lower edge does; kcand was built from the lower edges (hp_hiccups_score)
                    cand |= (ci <= mc) && (obs >= inff[fl].z);
                } else if (memf[fl]) {
                    chk[fl] = (unsigned)ci;
                    const int4 inf = inff[fl];
                    const int kb = obs < inf.y - 1 ? obs : inf.y - 1;
                    if (pi < A.sh_pairs && ci <= kShI && kb < kShK)
                        smem_red_add(&sh.hist[((pi * 2 + fl) * kShI + (ci - 1)) * kShK + kb], 1u);
                    else
                        atomicAdd(&A.hist[(size_t)(pi * 2 + fl) * A.total_bins + inf.x + kb], 1u);
                    cand |= (obs >= inf.z);
                }
            }
        }
    }
    // warp-converged bookkeeping: valid counts per (pair, background), candidate staging
    if (NPW != 1) {
        const int npw = NPW > 0 ? NPW : A.npw;
#pragma unroll
        for (int fl = 0; fl < 2; ++fl) {
            const bool v = (flags & (fl ? HP_SF_VALID_Y : HP_SF_VALID_K)) != 0;
            for (int k = 0; k < npw; ++k) {
                const unsigned mv = __ballot_sync(0xffffffffu, v && pi == k);
                if (mv && lane == 0) smem_red_add(&sh.nval[k * 2 + fl], (unsigned)__popc(mv));
            }
        }
    }
    const unsigned mcand = __ballot_sync(0xffffffffu, cand);
    if (mcand) {
        unsigned base = 0;
        if (lane == 0) base = smem_atom_add(sh.cnt, (unsigned)__popc(mcand));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) {
            const unsigned slot = base + __popc(mcand & ((1u << lane) - 1u));
            Cand cr;
            cr.r = r; cr.d = d; cr.obs = obs;
            cr.pair = (unsigned char)pi; cr.flags = (unsigned char)flags; cr.chunk_k = (unsigned char)chk[0]; cr.chunk_y = (unsigned char)chk[1];
            cr.e_k = Ev[0]; cr.e_y = Ev[1];
            if (slot < kStage) {
                sh.stage[slot] = cr;
            } else {  // staging full: straight to the global list
                const unsigned g = atomicAdd(&A.cand_count[0], 1u);
                if (g < A.cand_cap) A.cand[g] = cr; else atomicAdd(&A.cand_count[1], 1u);
            }
        }
    }
}
// merge the per-thread totals of a single-pair kernel into the CTA totals (before score_epilogue)
__device__ __forceinline__ void tail_acc_flush(const ScoreSmem& sh, const TailAcc& acc) {
#pragma unroll
    for (int fl = 0; fl < 2; ++fl) {
        const unsigned nv = __reduce_add_sync(0xffffffffu, acc.nval[fl]);
        unsigned hi = (unsigned)(acc.emax[fl] >> 32);
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned lo = hi == mhi ? (unsigned)acc.emax[fl] : 0u;
        const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
        if ((threadIdx.x & 31) == 0) {
            if (nv) smem_red_add(&sh.nval[fl], nv);
            const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
            if (m) smem_red_max(&sh.emax[fl], m);
        }
    }
}

// the tile arrives in kTileParts boxes of consecutive planes, each on its own mbarrier and the far planes first: the
// column blocks the warps claim first (largest d) can start as soon as their planes have landed
constexpr int kTileParts = 4;
__device__ __forceinline__ void score_issue_tile(const ScoreSmem& sh, const CUtensorMap* tm, int BD, int NQ, int q0, int plane0) {
    if (threadIdx.x == 0) {
        const int per = (BD + kTileParts - 1) / kTileParts;        // the tensor map's box holds `per` planes
        for (int k = 0; k < kTileParts; ++k) mbar_init(sh.bar + k, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int k = kTileParts - 1; k >= 0; --k) {
            mbar_expect_tx(sh.bar + k, (uint32_t)(per * 4 * NQ * 8));
            tma_load_3d(sh.tile + (size_t)k * per * 4 * NQ, tm, q0, 0, plane0 + k * per, sh.bar + k);
        }
        *sh.cnt = 0;
        *sh.next = 0;
    }
}
// wait for the planes [lo, hi] (tile-relative) of the tile
__device__ __forceinline__ void score_wait_planes(const ScoreSmem& sh, int BD, int lo, int hi) {
    const int per = (BD + kTileParts - 1) / kTileParts;
    for (int k = lo / per; k <= hi / per && k < kTileParts; ++k) mbar_wait(sh.bar + k, 0);
}

template <bool ISSUE>
__device__ __forceinline__ void score_prologue(const ScoreArgs& A, const ScoreSmem& sh, const CUtensorMap* tm, int tile_bytes, int q0,
                                               int plane0, int sh_bins) {
    const Chunks& C = A.tab->chunks;
    if (ISSUE && threadIdx.x == 0) {
        mbar_init(sh.bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(sh.bar, (uint32_t)tile_bytes);
        tma_load_3d(sh.tile, tm, q0, 0, plane0, sh.bar);
        *sh.cnt = 0;
        *sh.next = 0;
    }
    for (int i = threadIdx.x; i < sh_bins; i += blockDim.x) sh.hist[i] = 0;
    if (threadIdx.x < 16) { sh.emax[threadIdx.x] = 0ull; sh.nval[threadIdx.x] = 0u; }
    for (int i = threadIdx.x; i < kChunkTab; i += blockDim.x) {
        const bool in = i <= C.maxchunk;
        sh.rv[i] = in ? C.rv[i] : INFINITY;
        sh.cinfo[i] = in ? make_int4(C.hoff[i], C.hw[i], C.kcand[i], 0) : make_int4(0, 1, 0x7fffffff, 0);
    }
    __syncthreads();
    if (ISSUE) mbar_wait(sh.bar, 0);
}

// flush: privatised histogram, counters, staged candidates
__device__ __forceinline__ void score_epilogue(const ScoreArgs& A, const ScoreSmem& sh, int sh_bins) {
    __syncthreads();
    const int total_bins = A.total_bins;
    for (int i = threadIdx.x; i < sh_bins; i += blockDim.x) {
        const unsigned v = sh.hist[i];
        if (v) {
            const int kb = i % kShK, ci = (i / kShK) % kShI + 1, lf = i / (kShK * kShI);
            atomicAdd(&A.hist[(size_t)lf * total_bins + sh.cinfo[ci].x + kb], v);
        }
    }
    if (threadIdx.x < 2 * A.npw) {
        if (sh.nval[threadIdx.x]) atomicAdd(&A.nvalid[threadIdx.x], (unsigned long long)sh.nval[threadIdx.x]);
        if (sh.emax[threadIdx.x]) atomicMax(&A.emax_bits[threadIdx.x], sh.emax[threadIdx.x]);
    }
    unsigned staged = *sh.cnt;
    if (staged > kStage) staged = kStage;
    if (staged) {
        if (threadIdx.x == 0) *sh.base = atomicAdd(&A.cand_count[0], staged);
        __syncthreads();
        const unsigned base = *sh.base;
        for (unsigned i = threadIdx.x; i < staged; i += blockDim.x) {
            if (base + i < A.cand_cap) A.cand[base + i] = sh.stage[i];
            else atomicAdd(&A.cand_count[1], 1u);
        }
    }
}

// generic kernel: walks the op table of the sweep program, one pixel per thread per diagonal
__global__ void __launch_bounds__(kThreads) k_score(const __grid_constant__ CUtensorMap tm_bal, ScoreArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ScoreSmem sh = score_smem(smem, A.BD, A.NQ, 0, 0, 0, 0);
    const int sh_bins = A.sh_pairs * 2 * kShI * kShK;
    const int r0 = blockIdx.x * kTR;
    const int d0 = A.dlo + blockIdx.y * A.TD;
    const Tables& T = *A.tab;
    const int npw = T.prog.npw;
    score_prologue<true>(A, sh, &tm_bal, A.BD * 4 * A.NQ * 8, (r0 - A.HR) / 4, d0 - 2 * A.F, sh_bins);

    const int lane = threadIdx.x & 31;
    const int rl = threadIdx.x & (kTR - 1);
    const int r = r0 + rl;
    const int rho = rl + A.HR;                       // row inside the tile
    TailAcc tacc{};                                  // unused by the multi-pair tail

    for (int dl = threadIdx.x >> 7; dl < A.TD; dl += kThreads / kTR) {
        const int d = d0 + dl;                       // warp-uniform
        if (d > A.dhi) break;
        const bool inside = (r < A.n) && (r + d < A.n);
        unsigned char lv = kLvlNone;
        if (inside) lv = A.lvl[qidx(d, r, A.pitch)];
        int last = -1;
        if (lv < kLvlNever) {
            for (int pi = 0; pi < npw; ++pi) {
                if (d < T.prog.ww[pi]) continue;
                const int rs = T.prog.next_step[pi][lv];
                if (rs != kNoStep && rs > last) last = rs;
            }
        }
        const int wlast = __reduce_max_sync(0xffffffffu, last);
        if (wlast < 0) continue;
        double SK = 0.0, SY = 0.0;
        int opi = 0;
        for (int s = 0; s <= wlast; ++s) {
            const int oe = T.prog.op_end[s];
            if (s <= last) {
                for (; opi < oe; ++opi) {
                    const int a = T.opa[opi], b = T.opb[opi];
                    const int rr = rho + a;
                    const double v = sh.tile[((dl + 2 * A.F + b - a) * 4 + (rr & 3)) * A.NQ + (rr >> 2)];
                    SK = __dadd_rn(SK, v);
                    if (T.opy[opi]) SY = __dadd_rn(SY, v);
                }
            }
            const int pi = T.prog.step_pi[s];
            const bool em = (s <= last) && (d >= T.prog.ww[pi]) && (T.prog.next_step[pi][lv] == s);
            if (__any_sync(0xffffffffu, em)) emit_record<0, false>(A, sh, tacc, em, SK, SY, r, d, s, pi, lane, 0);
        }
    }
    score_epilogue(A, sh, sh_bins);
}

// ============================================================================================
// upload helper: plain diagonal-major staging plane -> quad-interleaved balanced plane, plus the
// "row has a non-zero stored balanced value" flags behind the gap mask (callers.py:238)
// ============================================================================================
template <typename T>
__global__ void k_relayout(const T* __restrict__ src, T* __restrict__ dst, unsigned int* __restrict__ rownz, int pitch, int num,
                           unsigned int* __restrict__ domain_bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (r >= pitch || d >= num) return;
    const T v = src[(size_t)d * pitch + r];
    dst[qidx(d, r, pitch)] = v;
    if (rownz && v != T(0)) rownz[r] = 1u;
    if constexpr (std::is_same<T, double>::value) {
        if (domain_bad && fast_domain_bad(v)) *domain_bad = 1u;
    }
}

// ============================================================================================
// Poisson tables: p[i][k] = 1 - pdtr(k, rv_i)  (callers.py:268-270); universal, built once per ctx
// ============================================================================================
__global__ void k_ptab(const Tables* __restrict__ tab, double* __restrict__ ptab) {
    const Chunks& C = tab->chunks;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C.total_bins) return;
    int lo = 1, hi = C.maxchunk;          // chunk containing flat bin idx
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (C.hoff[mid] <= idx) lo = mid; else hi = mid - 1;
    }
    ptab[idx] = poisson_sf((double)(idx - C.hoff[lo]), C.rv[lo]);
}

// ============================================================================================
// K3  Benjamini-Hochberg per lambda-chunk from the histogram -- replaces multipletests(fdr_bh)
// (callers.py:273): with N>=(k) = #pixels of the chunk with observed >= k and n the chunk size,
// raw(k) = p(k) / (N>=(k) / float(n)),  q(k) = min(1, min_{k' <= k, h[k'] > 0} raw(k')).
// One CTA per (chunk, pair*2+background).
// ============================================================================================
__device__ __forceinline__ unsigned long long block_scan_add(unsigned long long v, unsigned long long* sh, unsigned long long& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    __syncthreads();
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    unsigned long long pre = 0, tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) { if (w < wid) pre += sh[w]; tot += sh[w]; }
    total = tot;
    return x + pre;   // inclusive
}
__device__ __forceinline__ double block_scan_min(double v, double* sh, double& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = fmin(x, t);
    }
    __syncthreads();
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    double pre = INFINITY, tot = INFINITY;
    for (int w = 0; w < kThreads / 32; ++w) { if (w < wid) pre = fmin(pre, sh[w]); tot = fmin(tot, sh[w]); }
    total = tot;
    return fmin(x, pre);
}

__global__ void __launch_bounds__(kThreads) k_bh(const Tables* __restrict__ tab, const unsigned int* __restrict__ hist,
                                                  const double* __restrict__ ptab, double* __restrict__ qtab,
                                                  const int* __restrict__ numbin, int* __restrict__ kq, double sig) {
    const Chunks& c_chunks = tab->chunks;
    __shared__ unsigned long long sh_u[kThreads / 32];
    __shared__ double sh_d[kThreads / 32];
    const int ci = blockIdx.x + 1, lf = blockIdx.y;
    if (ci > numbin[lf]) return;
    const int w = c_chunks.hw[ci], off = c_chunks.hoff[ci];
    const unsigned int* h = hist + (size_t)lf * c_chunks.total_bins + off;
    double* q = qtab + (size_t)lf * c_chunks.total_bins + off;
    const double* p = ptab + off;
    // n = chunk size
    unsigned long long part = 0, n = 0;
    for (int k = threadIdx.x; k < w; k += kThreads) part += h[k];
    block_scan_add(part, sh_u, n);
    __syncthreads();
    if (n == 0) {
        for (int k = threadIdx.x; k < w; k += kThreads) q[k] = 1.0;
        if (kq && threadIdx.x == 0 && 1.0 <= sig) kq[lf * (kMaxChunk + 2) + ci] = 0;
        return;
    }
    const double fn = (double)n;
    // pass 1 (k descending): suffix counts -> raw(k) stored in q
    unsigned long long carry = 0;
    for (int base = w - 1; base >= 0; base -= kThreads) {
        const int k = base - (int)threadIdx.x;
        const unsigned long long hv = (k >= 0) ? h[k] : 0ull;
        unsigned long long tot;
        const unsigned long long inc = block_scan_add(hv, sh_u, tot) + carry;
        if (k >= 0) q[k] = hv ? __ddiv_rn(p[k], __ddiv_rn((double)inc, fn)) : INFINITY;
        carry += tot;
        __syncthreads();
    }
    __syncthreads();
    // pass 2 (k ascending): running minimum, clip at 1
    double cmin = INFINITY;
    for (int base = 0; base < w; base += kThreads) {
        const int k = base + (int)threadIdx.x;
        const double rv = (k < w) ? q[k] : INFINITY;
        double tot;
        const double m = fmin(block_scan_min(rv, sh_d, tot), cmin);
        if (k < w) {
            const double qk = m > 1.0 ? 1.0 : m;
            q[k] = qk;
            // q never rises with k: the first count with q <= sig is the chunk's survivor threshold (k_filter_fast)
            if (kq && qk <= sig) atomicMin(&kq[lf * (kMaxChunk + 2) + ci], k);
        }
        cmin = fmin(cmin, tot);
        __syncthreads();
    }
}

// ============================================================================================
// survivor selection: q <= sig for K or Y (callers.py:279-287), over the candidate list only
// ============================================================================================
struct FilterArgs {
    const Tables* tab;
    const Cand* cand;
    unsigned int ncand;
    const double* ptab;
    const double* qtab;
    const int* numbin;              // [npw*2]
    const double* bal;
    hp_survivor* out;
    unsigned int* out_count;        // [0] survivors, [1] dropped
    unsigned int out_cap;
    unsigned long long* nreject;    // [npw*2]
    double sig;
    int pitch;
    int bhfdr;
};

// The counters are bumped once per warp (ballot + popc), not once per record: tens of thousands of atomics on one
// address serialise in L2 and used to be most of this kernel's time.
__global__ void k_filter(FilterArgs A) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < A.ncand;
    const unsigned lane = threadIdx.x & 31u;
    const Chunks& c_chunks = A.tab->chunks;
    Cand c{};
    if (live) c = A.cand[idx];
    hp_survivor sv;
    sv.r = c.r; sv.c = c.r + c.d; sv.pair = c.pair; sv.flags = c.flags;
    sv.obs = (double)c.obs;
    sv.e[0] = c.e_k; sv.e[1] = c.e_y;
    bool rej[2] = {false, false};
    if (A.bhfdr) {
        // per-pixel Poisson rate (callers.py:536-540); BH over the whole chromosome is finished by the caller on the
        // pixels with p <= sig (every rejected pixel has p <= sig), q is filled in there
        const double p = live ? poisson_sf((double)c.obs, c.e_k) : 1.0;
        sv.p[0] = p; sv.q[0] = 1.0; sv.p[1] = 1.0; sv.q[1] = 1.0;
        rej[0] = live && p <= A.sig * (1.0 + 1e-9);
    } else {
#pragma unroll
        for (int fl = 0; fl < 2; ++fl) {
            const int lf = c.pair * 2 + fl;
            const int ci = fl ? c.chunk_y : c.chunk_k;
            double p = 1.0, q = 1.0;
            if (live && ci >= 1 && ci <= A.numbin[lf]) {
                const int w = c_chunks.hw[ci];
                const int kb = c.obs < w - 1 ? c.obs : w - 1;
                p = A.ptab[c_chunks.hoff[ci] + kb];
                q = A.qtab[(size_t)lf * c_chunks.total_bins + c_chunks.hoff[ci] + kb];
            }
            sv.p[fl] = p; sv.q[fl] = q;
            const bool valid = (c.flags & (fl ? HP_SF_VALID_Y : HP_SF_VALID_K)) != 0;
            rej[fl] = live && valid && q <= A.sig;
        }
    }
    if (rej[0]) sv.flags |= HP_SF_REJECT_K;
    if (rej[1]) sv.flags |= HP_SF_REJECT_Y;
    // rejected counts per (pair, background): lanes of a warp that share a pair add once
#pragma unroll
    for (int fl = 0; fl < 2; ++fl) {
        const unsigned peers = __match_any_sync(0xffffffffu, rej[fl] ? (int)c.pair : -1);
        if (rej[fl] && lane == (unsigned)__ffs(peers) - 1u) atomicAdd(&A.nreject[c.pair * 2 + fl], (unsigned long long)__popc(peers));
    }
    const bool any = rej[0] || rej[1];
    const unsigned many = __ballot_sync(0xffffffffu, any);
    if (many) {
        unsigned base = 0;
        if (lane == (unsigned)__ffs(many) - 1u) base = atomicAdd(&A.out_count[0], (unsigned)__popc(many));
        base = __shfl_sync(0xffffffffu, base, __ffs(many) - 1);
        if (any) {
            sv.ice = A.bal[qidx(c.d, c.r, A.pitch)];
            const unsigned g = base + __popc(many & ((1u << lane) - 1u));
            if (g < A.out_cap) A.out[g] = sv; else atomicAdd(&A.out_count[1], 1u);
        }
    }
}

__global__ void k_poisson_sf(const double* __restrict__ k, const double* __restrict__ mu, double* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = poisson_sf(k[i], mu[i]);
}

__global__ void k_fill_f64(double* __restrict__ p, double v, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace hp
