// k_score_spec: the score kernel for sweep programs known at compile time.
//
// The reference accumulates every donut cell into one fp64 sum per pixel in a fixed order (steps in
// (w, p) order; inside a step rows a = -w..w outer, columns b = -w..w inner, callers.py:147-198), and the
// sums must be reproduced bit for bit because lambda-chunk membership and the fold thresholds are
// discontinuous in them.  That order leaves no algebraic reuse between pixels, so the kernel is
// bound by fp64 add issue and by shared-memory operand bandwidth, not by HBM.  This kernel attacks
// both: the whole program (which rings every step adds) is a template parameter, so each add is one
// DADD fed by one LDS.64 with an immediate offset, and every thread owns a 4 x kTC block of pixels
// (4 consecutive matrix rows x kTC = 2 consecutive columns) so that a loaded operand feeds up to
// 4 x kTC accumulators from registers (0.26 loads per add; 4 x 4 blocks need 200+ registers and ran slower).  The quad-interleaved plane layout (hp_device.cuh)
// makes those loads conflict-free: lane l owns rows 4l..4l+3, and for a fixed row offset the 32
// lanes read 32 consecutive doubles.
#pragma once
#include <type_traits>

#include "hp_kernels.cuh"

namespace hp {

// ---- compile-time sweep program (callers.py:15-23 step order, :150-152 skip rule) -------------
template <int MAXW_, int NPW_, int P0, int W0, int P1 = 0, int W1 = 0, int P2 = 0, int W2 = 0, int P3 = 0, int W3 = 0>
struct SProg {
    static constexpr int maxww = MAXW_;
    static constexpr int npw = NPW_;
    __host__ __device__ static constexpr int pw(int i) { return i == 0 ? P0 : i == 1 ? P1 : i == 2 ? P2 : P3; }
    __host__ __device__ static constexpr int ww(int i) { return i == 0 ? W0 : i == 1 ? W1 : i == 2 ? W2 : W3; }
    __host__ __device__ static constexpr int nsteps() {
        int k = 0;
        for (int i = 0; i < npw; ++i) k += maxww - ww(i) + 1;
        return k;
    }
    // pair index of step s; steps sorted by (w, p)
    __host__ __device__ static constexpr int step_pi(int s) {
        int k = 0;
        for (int w = 1; w <= maxww; ++w) {
            int lastp = -1;                        // pairs with ww <= w in ascending p
            for (int t = 0; t < npw; ++t) {
                int best = -1;
                for (int i = 0; i < npw; ++i)
                    if (ww(i) <= w && pw(i) > lastp && (best < 0 || pw(i) < pw(best))) best = i;
                if (best < 0) break;
                if (k == s) return best;
                ++k;
                lastp = pw(best);
            }
        }
        return -1;
    }
    __host__ __device__ static constexpr int step_w(int s) {
        int k = 0;
        for (int w = 1; w <= maxww; ++w)
            for (int i = 0; i < npw; ++i)
                if (ww(i) <= w) { if (k == s) return w; ++k; }
        return -1;
    }
    __host__ __device__ static constexpr int step_p(int s) { return pw(step_pi(s)); }
    // rings g = max(|a|, |b|) whose cells (off the cross) step s adds
    __host__ __device__ static constexpr unsigned mask(int s) {
        const int p = step_p(s), w = step_w(s);
        const bool limit = s > 0;
        const int lp = limit ? step_p(s - 1) : 0, lw = limit ? step_w(s - 1) : 0;
        const int mx = p > lp ? p : lp, mn = p < lp ? p : lp;
        unsigned m = 0;
        for (int g = 1; g <= w; ++g) {
            if (limit && ((g <= lw && g > mx) || g <= mn)) continue;
            if (g <= p) continue;                  // current peak square
            m |= 1u << g;
        }
        return m;
    }
    __host__ __device__ static constexpr int min_pw() {
        int m = pw(0);
        for (int i = 1; i < npw; ++i) m = pw(i) < m ? pw(i) : m;
        return m;
    }
    // rings whose lower-left cells step s also adds to Reads (callers.py:197-198: the first step, then
    // only steps of the smallest p and only rings beyond the previous step's w)
    __host__ __device__ static constexpr unsigned rmask(int s) {
        const unsigned m = mask(s);
        if (s == 0) return m;
        if (step_p(s) != min_pw()) return 0u;
        unsigned out = 0;
        for (int g = step_w(s - 1) + 1; g < 32; ++g) out |= m & (1u << g);
        return out;
    }
    // previous step of the same pair, -1 if none: pixels with level in (prev, s] resolve the pair at s
    __host__ __device__ static constexpr int prev_same_pair(int s) {
        const int pi = step_pi(s);
        for (int t = s - 1; t >= 0; --t)
            if (step_pi(t) == pi) return t;
        return -1;
    }
};

__host__ __device__ constexpr int hp_top_ring(unsigned m) {
    int w = 0;
    for (int g = 1; g < 32; ++g) if ((m >> g) & 1u) w = g;
    return w;
}

constexpr int kTC = 2;              // pixel block: 4 rows x kTC columns per thread
constexpr int kSpecThreads = 512;   // 16 warps share one tile (one CTA per SM)
constexpr int kQCap = 32 + kTC * 32;   // per-warp tail queue: < 32 left over + one pixel row of every lane's block

// offset (in doubles) of matrix element (row r + rho, column c + kappa) from the thread's base
// pointer &tile[d'(c - r)][0][quad(r)]
__host__ __device__ constexpr int spec_off(int rho, int kappa) {
    return ((kappa - rho) * 4 + (rho & 3)) * kNQ + (rho >> 2);
}

// compile-time loop: f(integral_constant<int, B>), ..., f(integral_constant<int, E - 1>).  Used instead of
// `#pragma unroll` because every condition inside the sweep must fold to a constant (nvcc keeps the
// big outer loops rolled and predicates the adds otherwise).
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

__host__ __device__ constexpr int hp_iabs(int x) { return x < 0 ? -x : x; }
__host__ __device__ constexpr bool hp_cell_in(unsigned mask, int a, int b) {
    return a != 0 && b != 0 && ((mask >> (hp_iabs(a) > hp_iabs(b) ? hp_iabs(a) : hp_iabs(b))) & 1u) != 0u;
}
// does matrix element (rho, kappa) of the block's window feed any of the 16 pixels under MASK / W?
__host__ __device__ constexpr bool hp_strip_used(unsigned mask, int W, int rho, int kappa) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < kTC; ++j) {
            const int a = rho - i, b = kappa - j;
            if (a >= -W && a <= W && b >= -W && b <= W && hp_cell_in(mask, a, b)) return true;
        }
    return false;
}

// one sweep step: add the cells of the rings in MASK to the 16 pixels of the thread
template <unsigned MASK>
__device__ __forceinline__ void spec_accumulate(const double* __restrict__ base, double (&aK)[4][kTC], double (&aY)[4][kTC]) {
    constexpr int W = hp_top_ring(MASK);
    static_for<-W, W + 4>([&](auto RHO) {                 // matrix row, relative to the block's first row
        constexpr int rho = decltype(RHO)::value;
        double strip[2 * W + kTC];
        static_for<0, 2 * W + kTC>([&](auto KK) {
            constexpr int k = decltype(KK)::value;
            if constexpr (hp_strip_used(MASK, W, rho, k - W)) strip[k] = base[spec_off(rho, k - W)];
        });
        static_for<0, 4>([&](auto II) {
            constexpr int i = decltype(II)::value;
            constexpr int a = rho - i;
            if constexpr (a >= -W && a <= W && a != 0) {
                static_for<0, kTC>([&](auto JJ) {
                    constexpr int j = decltype(JJ)::value;
                    static_for<-W, W + 1>([&](auto BB) {
                        constexpr int b = decltype(BB)::value;
                        if constexpr (hp_cell_in(MASK, a, b)) {
                            const double v = strip[j + b + W];
                            aK[i][j] = __dadd_rn(aK[i][j], v);
                            if constexpr (a > 0 && b < 0) aY[i][j] = __dadd_rn(aY[i][j], v);
                        }
                    });
                });
            }
        });
    });
}

// run-time step index -> the compile-time accumulate code of that step
template <class PG, int S>
__device__ __forceinline__ void spec_dispatch(int s, const double* base, double (&aK)[4][kTC], double (&aY)[4][kTC]) {
    if constexpr (S < PG::nsteps()) {
        if (s == S) {
            constexpr unsigned M = PG::mask(S);
            if constexpr (M != 0u) spec_accumulate<M>(base, aK, aY);
        } else {
            spec_dispatch<PG, S + 1>(s, base, aK, aY);
        }
    }
}

// CTA: kTR rows x TD diagonals of (row, column)-aligned 4 x kTC pixel blocks; a warp claims the column block
// [dc, dc + kTC), dc = d0 + kTC * kb, dynamically (deepest first), lane l owns the rows 4l..4l+3.  Rows with (r & 3) = i
// cover the diagonals [d0 - i, d0 + TD - i): consecutive CTAs in d tile the band without overlap,
// and the first / last CTA mask the <= 3 diagonals that stick out of [dlo, dhi].
//
// Per step the warp accumulates (all lanes, all 16 pixels: zero pixels ride along), then pixels that
// resolve a pair at this step push (bS_K, bS_Y, r, d, step, pair) on a per-warp queue; whenever 32
// records are waiting the warp runs the per-pixel tail on them fully converged.
#ifndef HP_SPEC_MINB
#define HP_SPEC_MINB 1
#endif
template <class PG>
__global__ void __launch_bounds__(kSpecThreads, HP_SPEC_MINB) k_score_spec(const __grid_constant__ CUtensorMap tm_bal,
                                                                            const __grid_constant__ ScoreArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    ScoreSmem sh = score_smem(smem, A.BD, kNQ, kSpecThreads / 32, kQCap, A.nexec, A.TD + 8);
    const int sh_bins = A.sh_pairs * 2 * kShI * kShK;
    const int r0 = blockIdx.x * kTR;
    // far diagonals first: they run the deepest levels, so the cheap tiles fill the tail of the grid
    const int d0 = A.dlo + (int)(gridDim.y - 1 - blockIdx.y) * A.TD;
    const int plane0 = d0 - 3 - 2 * A.F;
    score_issue_tile(sh, &tm_bal, A.BD, kNQ, (r0 - kHR) / 4, plane0);      // TMA first; the table copies below overlap it
    // per-CTA copies of the tables the tail reads: bE of interior pixels, IR, biases of the tile's rows / columns
    sh.tb_d0 = d0 - 3; sh.tb_r0 = r0;
    {
        const int nd = sh.tb_nd, nexec = A.nexec;
        for (int t = threadIdx.x; t < 2 * nexec * nd; t += kSpecThreads) {
            const int dd = t % nd, fs = t / nd, d = sh.tb_d0 + dd;       // fs = fl * nexec + s; interior table is z = 0
            sh.tb_be[t] = (d >= 0 && d < A.num) ? A.betab[(size_t)fs * A.num + d] : 0.0;
        }
        for (int t = threadIdx.x; t < nd; t += kSpecThreads) {
            const int d = sh.tb_d0 + t;
            sh.tb_ir[t] = (d >= 0 && d < A.num) ? A.ir[d] : 0.0;
        }
        for (int t = threadIdx.x; t < kTR; t += kSpecThreads) sh.tb_b1[t] = r0 + t < A.n ? A.b1[r0 + t] : 0.0;
        for (int t = threadIdx.x; t < kTR + nd; t += kSpecThreads) {
            const int c = r0 + sh.tb_d0 + t;
            sh.tb_b2[t] = (c >= 0 && c < A.n) ? A.b2[c] : 0.0;
        }
    }
    score_prologue<false>(A, sh, &tm_bal, 0, 0, 0, sh_bins);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = r0 + 4 * lane;
    const unsigned lt = (1u << lane) - 1u;
    const int dspan = A.dspan;
    double2* qs = sh.qsum + warp * kQCap;
    int2* qm = sh.qmeta + warp * kQCap;
    int* qo = sh.qobs + warp * kQCap;
    int cnt = 0;                                     // warp-uniform queue fill
    TailAcc tacc{};
    const int ntask = min(A.TD / kTC, (A.dhi + 3 - d0) / kTC + 1);   // column blocks with a pixel in [dlo, dhi]
    for (;;) {
        int kb = 0;                                  // column blocks are claimed dynamically, deepest (largest d) first
#ifdef HP_TASK_ASC
        if (lane == 0) { kb = (int)smem_atom_add(sh.next, 1u); if (kb >= ntask) kb = -1; }
#else
        if (lane == 0) kb = ntask - 1 - (int)smem_atom_add(sh.next, 1u);
#endif
        kb = __shfl_sync(0xffffffffu, kb, 0);
        if (kb < 0) break;
        const int dc = d0 + kb * kTC;                // diagonal of pixel (0, 0) of the block; warp-uniform
        unsigned lvp[4];                             // levels of the 16 pixels, one byte each, [i] = row
        int last = -1;
        // whole block inside the band and the chromosome (true for almost every block): no per-pixel bounds checks
        const bool interior = dc - 3 >= A.dlo && dc + kTC - 1 <= A.dhi && r0 + kTR - 1 + dc + kTC - 1 < A.n;
        int obsv[4][kTC];                            // raw counts, loaded now so that the tail never waits on HBM
        if (interior) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned pk = 0;
#pragma unroll
                for (int j = 0; j < kTC; ++j) {
                    const size_t q = qidx(dc + j - i, r + i, A.pitch);
                    pk |= (unsigned)A.lvl[q] << (8 * j);
                    obsv[i][j] = A.raw[q];
                }
                lvp[i] = pk;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned pk = 0;
#pragma unroll
                for (int j = 0; j < kTC; ++j) {
                    const int rr = r + i, d = dc + j - i;
                    unsigned lv = kLvlNone;
                    obsv[i][j] = 0;
                    if (d >= A.dlo && d <= A.dhi && rr < A.n && rr + d < A.n) {
                        const size_t q = qidx(d, rr, A.pitch);
                        lv = A.lvl[q];
                        obsv[i][j] = A.raw[q];
                    }
                    pk |= lv << (8 * j);
                }
                lvp[i] = pk;
            }
        }
        // step 0 does not depend on the levels: run it while the level / count loads above are still in flight
        if (r0 + dc - 3 >= A.n) continue;                            // every pixel of the block is beyond the chromosome
        double aK[4][kTC], aY[4][kTC];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < kTC; ++j) { aK[i][j] = 0.0; aY[i][j] = 0.0; }
        score_wait_planes(sh, A.BD, dc - plane0 - 3 - 2 * A.F, dc - plane0 + kTC + 2 + 2 * A.F);
        const double* base = sh.tile + (size_t)(dc - plane0) * 4 * kNQ + (lane + kHR / 4);
        spec_dispatch<PG, 0>(0, base, aK, aY);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < kTC; ++j) {
                const unsigned lv = (lvp[i] >> (8 * j)) & 0xffu;
                if (lv < kLvlNever) {
                    const int d = dc + j - i;
                    const int k = d - A.dlo < dspan ? d - A.dlo : dspan;
                    const int rs = A.last_need[k][lv];
                    if (rs != kNoStep && rs > last) last = rs;
                }
            }
        const int wlast = __reduce_max_sync(0xffffffffu, last);
        if (wlast < 0) continue;
#pragma unroll 1
        for (int s = 0; s <= wlast; ++s) {
            if (s > 0) spec_dispatch<PG, 0>(s, base, aK, aY);
            // pixels whose level lies in (previous executed step of this pair, s] resolve the pair now
            const int pi = A.step_pi[s], lo = A.step_lo[s], wmin = A.ww[pi];
            unsigned em = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < kTC; ++j) {
                    const int lv = (lvp[i] >> (8 * j)) & 0xff;
                    // single pair: level s resolves at step s (every in-band diagonal is >= ww)
                    const bool e = PG::npw == 1 ? lv == s : (lv >= lo && lv <= s && (dc + j - i) >= wmin);
                    if (e) em |= 1u << (i * kTC + j);
                }
            if (!__any_sync(0xffffffffu, em != 0u)) continue;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const unsigned emr = (em >> (i * kTC)) & ((1u << kTC) - 1u);
                if (!__any_sync(0xffffffffu, emr != 0u)) continue;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
                    if (ii == i) {
#pragma unroll
                        for (int j = 0; j < kTC; ++j) {
                            const bool e = (emr >> j) & 1u;
                            const unsigned m = __ballot_sync(0xffffffffu, e);
                            if (e) {
                                const int slot = cnt + __popc(m & lt);
                                qs[slot] = make_double2(aK[ii][j], aY[ii][j]);
                                qm[slot] = make_int2(r + ii, ((dc + j - ii) << 16) | (s << 8) | pi);
                                qo[slot] = obsv[ii][j];
                            }
                            cnt += __popc(m);
                        }
                    }
                __syncwarp();
                while (cnt >= 32) {
                    cnt -= 32;
                    const double2 v = qs[cnt + lane];
                    const int2 mt = qm[cnt + lane];
                    emit_record<PG::npw, true>(A, sh, tacc, true, v.x, v.y, mt.x, mt.y >> 16, (mt.y >> 8) & 0xff, mt.y & 0xff, lane,
                                               qo[cnt + lane]);
                }
                __syncwarp();
            }
        }
    }
    for (; cnt > 0; cnt -= 32) {                      // leftover records of this warp
        const int k = cnt - 1 - lane;
        const bool act = k >= 0;
        const double2 v = act ? qs[k] : make_double2(0.0, 0.0);
        const int2 mt = act ? qm[k] : make_int2(0, 0);
        emit_record<PG::npw, true>(A, sh, tacc, act, v.x, v.y, mt.x, mt.y >> 16, (mt.y >> 8) & 0xff, mt.y & 0xff, lane,
                                   act ? qo[k] : 0);
    }
    if (PG::npw == 1) tail_acc_flush(sh, tacc);
    score_epilogue(A, sh, sh_bins);
}

// ============================================================================================
// k_levels_spec -- K1 for compiled-in sweep programs: the raw lower-left ring sums (integers, so any
// order) with the same 4-row register blocks, 4 columns per thread.  Output: per non-zero band pixel
// the first step s* with Reads >= min_local_reads (callers.py:203-206) and the histogram of s*.
// ============================================================================================
constexpr int kTCL = 4;

template <unsigned RMASK>
__device__ __forceinline__ void level_accumulate(const int* __restrict__ base, int (&R)[4][kTCL]) {
    constexpr int W = hp_top_ring(RMASK);
    static_for<1, W + 4>([&](auto RHO) {                   // rows below the block's first row
        constexpr int rho = decltype(RHO)::value;
        int strip[W + kTCL];                               // columns c - W .. c + kTCL - 1
        static_for<0, W + kTCL>([&](auto KK) {
            constexpr int k = decltype(KK)::value;
            strip[k] = base[((k - W - rho) * 4 + (rho & 3)) * kNQL + (rho >> 2)];
        });
        static_for<0, 4>([&](auto II) {
            constexpr int i = decltype(II)::value;
            constexpr int a = rho - i;
            if constexpr (a >= 1 && a <= W) {
                static_for<0, kTCL>([&](auto JJ) {
                    constexpr int j = decltype(JJ)::value;
                    static_for<-W, 0>([&](auto BB) {
                        constexpr int b = decltype(BB)::value;
                        if constexpr (hp_cell_in(RMASK, a, b)) R[i][j] += strip[j + b + W];
                    });
                });
            }
        });
    });
}

template <class PG, int S>
__device__ __forceinline__ void level_dispatch(int s, const int* base, int (&R)[4][kTCL]) {
    if constexpr (S < PG::nsteps()) {
        if (s == S) {
            constexpr unsigned M = PG::rmask(S);
            if constexpr (M != 0u) level_accumulate<M>(base, R);
        } else {
            level_dispatch<PG, S + 1>(s, base, R);
        }
    }
}

// tile: rows [r0, r0 + 4 kNQL), planes [d0 - 3 - 2F, d0 + TD); thread = rows 4l..4l+3 x 4 columns
template <class PG>
__global__ void __launch_bounds__(kThreads) k_levels_spec(const __grid_constant__ CUtensorMap tm_raw, LevelArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Tables& T = *A.tab;
    int* tile = reinterpret_cast<int*>(smem);
    const int tile_bytes = A.BD * 4 * kNQL * 4;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + tile_bytes);
    unsigned int* next = reinterpret_cast<unsigned int*>(smem + tile_bytes + 8);
    unsigned int* sh_hist = reinterpret_cast<unsigned int*>(smem + tile_bytes + 16);
    const int r0 = blockIdx.x * kTR;
    const int d0 = A.dlo + (int)(gridDim.y - 1 - blockIdx.y) * A.TD;
    const int plane0 = d0 - 3 - 2 * A.F;
    const int nsteps = T.prog.nsteps;
    for (int i = threadIdx.x; i <= nsteps; i += kThreads) sh_hist[i] = 0;
    if (threadIdx.x == 0) *next = 0;
    level_prologue(tile, bar, &tm_raw, tile_bytes, r0 / 4, plane0);

    const int lane = threadIdx.x & 31;
    const int r = r0 + 4 * lane;
    const int thr = T.prog.thr;
    for (;;) {
        int kb = 0;
        if (lane == 0) kb = (int)smem_atom_add(next, 1u);
        kb = __shfl_sync(0xffffffffu, kb, 0);
        const int dc = d0 + kb * kTCL;
        if (kb * kTCL >= A.TD || dc - 3 > A.dhi) break;
        const int* base = tile + (size_t)(dc - plane0) * 4 * kNQL + lane;
        int R[4][kTCL];
        unsigned lvp[4];                                   // one byte per pixel: kLvlNone / kLvlNever / step
        bool open = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned pk = 0;
#pragma unroll
            for (int j = 0; j < kTCL; ++j) {
                R[i][j] = 0;
                const int rr = r + i, d = dc + j - i;
                unsigned lv = kLvlNone;
                if (d >= A.dlo && d <= A.dhi && rr < A.n && rr + d < A.n && base[((j - i) * 4 + i) * kNQL] != 0) lv = kLvlNever;
                pk |= lv << (8 * j);
            }
            lvp[i] = pk;
            open |= __vcmpeq4(pk, 0xFEFEFEFEu) != 0u;
        }
        int never = 0;
        for (int s = 0; s < nsteps && __any_sync(0xffffffffu, open); ++s) {
            level_dispatch<PG, 0>(s, base, R);
            int hit = 0;
            open = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < kTCL; ++j) {
                    const unsigned sh8 = 8 * j;
                    if (((lvp[i] >> sh8) & 0xffu) == kLvlNever && R[i][j] >= thr) {
                        lvp[i] = (lvp[i] & ~(0xffu << sh8)) | ((unsigned)s << sh8);
                        ++hit;
                    }
                }
                open |= __vcmpeq4(lvp[i], 0xFEFEFEFEu) != 0u;
            }
            hit = __reduce_add_sync(0xffffffffu, hit);
            if (lane == 0 && hit) smem_red_add(&sh_hist[s], (unsigned)hit);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            never += __popc(__vcmpeq4(lvp[i], 0xFEFEFEFEu)) >> 3;
#pragma unroll
            for (int j = 0; j < kTCL; ++j) {
                const int rr = r + i, d = dc + j - i;
                if (d >= A.dlo && d <= A.dhi && rr < A.n) A.lvl[qidx(d, rr, A.pitch)] = (unsigned char)(lvp[i] >> (8 * j));
            }
        }
        never = __reduce_add_sync(0xffffffffu, never);
        if (lane == 0 && never) smem_red_add(&sh_hist[nsteps], (unsigned)never);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= nsteps; i += kThreads)
        if (sh_hist[i]) atomicAdd(&A.hist[i], (unsigned long long)sh_hist[i]);
}

// ---- host side: does a run-time program (a prefix of it) equal the compiled one? ----------------
template <class PG>
inline bool spec_matches(const Prog& G, int nexec, const signed char* opa, const signed char* opb, const unsigned char* opy,
                         const unsigned char* opr) {
    if (G.npw != PG::npw || nexec > PG::nsteps() || nexec > G.nsteps) return false;
    for (int i = 0; i < PG::npw; ++i)
        if (G.pw[i] != PG::pw(i) || G.ww[i] != PG::ww(i)) return false;
    int k = 0;
    for (int s = 0; s < nexec; ++s) {
        if (G.step_pi[s] != PG::step_pi(s) || G.step_w[s] != PG::step_w(s)) return false;
        const unsigned m = PG::mask(s);
        const int w = G.step_w[s];
        for (int a = -w; a <= w; ++a)
            for (int b = -w; b <= w; ++b) {
                if (a == 0 || b == 0) continue;
                const int g = abs(a) > abs(b) ? abs(a) : abs(b);
                if (!((m >> g) & 1u)) continue;
                if (k >= G.op_end[s] || opa[k] != a || opb[k] != b || opy[k] != (unsigned char)(a > 0 && b < 0)) return false;
                if (opr[k] != (unsigned char)(a > 0 && b < 0 && ((PG::rmask(s) >> g) & 1u))) return false;
                ++k;
            }
        if (k != G.op_end[s]) return false;
    }
    return true;
}

}  // namespace hp
