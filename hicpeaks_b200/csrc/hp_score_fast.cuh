// k_score_fast: the re-associated score kernel with guard-band exact re-evaluation (single (p, w) programs).
//
// What must be bit-exact in the reference's output is every DECISION made on the fp64 donut / lower-left sums
// (callers.py:132-198): is E > 0, which lambda-chunk does E fall in (strict edges, callers.py:38), is the pixel a
// candidate -- and the E of the few pixels that are finally reported.  The sums themselves are only ever seen through
// those.  So this kernel evaluates the sums in fp32 and in a convenient order, carries a RIGOROUS bound on the
// difference to the reference's fp64 value (all summands are >= 0, so every partial sum is bounded by the span sum and
// the rounding error of any summation order is bounded by ops * 2^-24 * span sum), and classifies each pixel only when
// the whole interval [lo, hi] around its expected value lies strictly inside one lambda-chunk.  A pixel whose interval
// touches a chunk edge or zero, a pixel next to both chromosome ends, and every pixel that could hold E.max() go to a
// short list that k_exact (hp_exact.cuh) re-evaluates in the reference's exact fp64 order; the survivors of the FDR step
// get their reported E the same way (k_fill_exact).  Histograms, counts, E.max(), survivors, E, p, q are therefore
// bit-identical to the exact kernels (k_score_spec / k_score); tests/test_gpu_fast.py holds the two paths against each
// other and against the oracle.  Inputs the bound does not cover (negative or non-finite balanced values, values
// outside [1e-30, 1e30]) raise a flag and the chromosome is re-run through the exact kernel.
//
// Arithmetic: for a pixel (r, c) and half-width w the four quadrant boxes are
//     Q_w = sum_{1<=|a|<=w} sum_{1<=|b|<=w} X[r+a, c+b],     LL_w = sum_{a=1..w} sum_{b=-w..-1} X[r+a, c+b]
// and the reference's donut / lower-left sums of a pixel resolving at width w are K = Q_w - Q_p, Y = LL_w - LL_p
// (cross and peak square excluded, callers.py:138-141).  A thread owns one matrix row and kFNPX = 8 consecutive
// columns; it keeps, for the 2 FM + 8 matrix columns its pixels can reach, the column sums over rows r+1..r+g (dn) and
// over rows r-g..r-1 plus r+1..r+g (W) in registers, extends them by one row pair per level g (two rows of 128-bit
// shared-memory loads: the tile is stored row-major so a matrix row is contiguous), and slides the (2w+1)-wide window
// over them at the levels some pixel of the warp resolves at.  ~110 fp32 adds per pixel instead of ~200 ordered fp64
// adds, 128-bit conflict-free loads, no barrier inside a tile's compute phase.
//
// Data movement: nothing is filled by threads.  The upload leaves an fp32 row-major copy of the balanced band in HBM
// (xf[r][d - min(ww)], k_f32plane) next to the fp64 planes the exact kernels read; a tile of it -- 96 matrix rows x PX
// diagonals, already in the order the sums want -- arrives by one TMA box, the raw counts and the levels of the tile and
// the three small factor tables by further TMA / bulk copies on the same mbarrier.  A CTA owns two such stages and a
// static list of tiles; its 16 warps claim the 16 (row half, column block) passes of tile after tile from one counter,
// wait only for the mbarrier of the stage they need, and the warp that finishes the last pass of a tile re-arms the stage
// with the tile after next: no CTA-wide barrier between the prologue and the epilogue, loads always one tile ahead.
#pragma once
#include "hp_kernels.cuh"

namespace hp {

constexpr int kFTR = 64;            // tile rows
constexpr int kFTD = 64;            // tile diagonals
constexpr int kFThreads = 512;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFNPX = 8;            // pixels (consecutive columns of one row) per thread and pass
constexpr int kFRowHalo = 16;       // tile row 0 is matrix row r0 - 16 (>= FM)
constexpr int kFXR = kFTR + 2 * kFRowHalo;
constexpr int kFStages = 2;         // tiles in shared memory per CTA
constexpr int kFPasses = 2 * (kFTD / kFNPX);    // (row half, column block) passes of a tile
constexpr int kFQCap = 32 * kFNPX;  // per-warp record queue: every pixel of a pass
constexpr unsigned kFScratch = 1024;            // E.max() contenders a CTA keeps until it knows its final lower bound
constexpr unsigned kFCandChunk = 128;           // candidate slots a warp reserves at a time (unused ones are marked r = -1)
constexpr float kFU = 5.9604645e-8f;           // 2^-24
constexpr float kFCRel = 16.f * kFU;           // relative error of the fp32 factor chain f * b1 * b2 * S
// |K' - K| <= fast_cerr_k(w) * (Fmax_w + Fmax_p), |Y' - Y| <= fast_cerr_y(w) * (LLmax_w + LLmax_p): Fmax / LLmax = the
// largest window sum among the thread's 8 pixels at that width (every intermediate of the sliding sums is below it).
// Derivation in DESIGN.md section 3 (inputs rounded once, 2w ordered adds per column sum, 2w adds for the first window,
// two roundings per slide, one per difference); 5 % on top for the second-order terms.
__host__ __device__ constexpr float fast_cerr_k(int w) { return (10.f * w + 17.f) * 1.05f * kFU; }
__host__ __device__ constexpr float fast_cerr_y(int w) { return (4.f * w + 16.f) * 1.05f * kFU; }
// General (multi-pair) form: K = sum_g c(g) Q_g with small integer coefficients (below); every Q'_g carries the bound
// above, every fused multiply-add of the chain (at most FM of them) rounds once more on a partial sum that is itself
// bounded by sum_g |c(g)| Fmax_g:  |K' - K| <= sum_g |c(g)| fast_cerr_gen_k(g) Fmax_g, likewise for Y.
__host__ __device__ constexpr float fast_cerr_gen_k(int w) { return (10.f * w + 40.f) * 1.05f * kFU; }
__host__ __device__ constexpr float fast_cerr_gen_y(int w) { return (4.f * w + 40.f) * 1.05f * kFU; }
constexpr int kFMaxPeak = 4;        // largest pw the single-pair kernel has window code for
constexpr int kFMinWidth = 3;       // smallest ww ...                                           (others: the general-form kernel)
constexpr int kFMaxCode = 16;       // widths a pair can resolve at (codes 0 .. maxww - ww), 0xF = none
constexpr int kFMaxG = 16;          // ring index bound of the coefficient tables (> FM)

struct XRec {                       // a record for k_exact: kind 0 = evaluate and account the whole record,
    int r, ds, obs, kind;           // kind bit 0 / 1 = only E.max() of K / Y needs the exact value.  ds = d | step << 16
};
struct FCand {                      // 16 B: a classified pixel whose Poisson p can pass sig (E is filled in later,
    int r, ds, obs;                 // for survivors only).  ds = d | step << 16
    unsigned info;                  // chunk_k | chunk_y << 8 | flags << 16 | pair << 24
};

struct FastArgs {
    const float* ffac;              // [1 + 2F][2][nexec][num]: fp32 IR[d] / bE, laid out like betab
                                    // (0: the pixel is certainly invalid, NaN: fp32 cannot hold the factor)
    const float* ffs;               // [nstrips][2][nexec][kFTD]: the interior (z = 0) part of ffac, one block per strip
    const float* b1f;               // fp32 biases of the rows, zero padded (k_fast_bias)
    const float* b2s;               // fp32 biases of the columns, shifted: b2s[i] = B2[i + dlo], zero padded
    const Tables* tab;
    unsigned int* hist;             // [2][total_bins]
    unsigned long long* nvalid;     // [2]
    FCand* fcand;
    XRec* xrec;
    int4* scratch;                  // [gridDim.x][kFScratch] (r, d | step << 16, hi of K or 0, hi of Y or 0)
    unsigned int* cnt;              // d_cnt: [2] chunk overflow, [8] fcand, [9] fcand dropped, [10] xrec, [11] xrec dropped
    unsigned int fcand_cap, xrec_cap;
    int n, num, pitch, dlo, dhi, F, nexec, maxchunk, total_bins;
    int nstrips, ntr;
    int p, w0;                      // the (pw, ww) pair of the run
    unsigned c1dn, c2dn, c1up, c2up;   // FastEdges (host: fast_edges())
    float elo[kMaxChunk + 4];       // [i] = lower edge of lambda-chunk i rounded up (with margin); [1] = 0
    unsigned int* gmax;             // [2] running max of lo over the CTAs (K, Y) for this pair
    // ---- general form (union programs: one launch per pair, GEN kernels) ------------------------------------------------
    // The reference keeps ONE set of accumulators for all pairs and re-adds rings whenever p drops back (callers.py:150-152),
    // so at executed step t the donut sum is sum_g m_t(g) Ring_g with multiplicities m_t(g) >= 0, i.e. sum_g c_t(g) Q_g with
    // c_t(g) = m_t(g) - m_t(g + 1).  A pixel resolves this pair at step t = next_step[pair][level]; code = width(t) - w0.
    int pair, ncode;                // pair index; codes of this pair among the executed steps
    int nlevels;                    // level codes below this are real steps (tcode has that many entries)
    int wpair;                      // ww of this pair: pixels on diagonals below it are not tested for the pair
    unsigned char tcode[HP_MAX_STEPS + 2];      // level -> code (0xF: the pair never resolves for that level)
    unsigned char tstep[kFMaxCode];             // code -> executed step index (the record's step)
    unsigned short hmask[kFMaxCode];            // code -> rings g with c(g) != 0
    float cabs[kFMaxG];                         // ring -> max over the codes of |c(g)|
    float ctab[kFMaxCode][kFMaxG];              // [code][g] = c(g)
};

__host__ __device__ constexpr int fast_px(int FM) {        // tile pitch in floats: >= 64 + 4 FM and == 4 (mod 8), so that
    int p = kFTD + 4 * FM;                                 // 32 lanes on consecutive rows read / write 128-bit words
    while (p % 8 != 4) ++p;                                // without bank conflicts
    return p;
}

template <int B, int E, class F>
__host__ __device__ __forceinline__ void sfor(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        sfor<B + 1, E>(f);
    }
}

__host__ __device__ __forceinline__ float fast_max(float a, float b) { return fmaxf(a, b); }

__host__ __device__ __forceinline__ unsigned fast_pack_err(float ek, float ey) {           // two bf16, rounded up
#ifdef __CUDA_ARCH__
    const unsigned a = __float_as_uint(ek), b = __float_as_uint(ey);
#else
    unsigned a, b; { union { float f; unsigned u; } x; x.f = ek; a = x.u; x.f = ey; b = x.u; }
#endif
    return ((a + 0xFFFFu) >> 16) | ((b + 0xFFFFu) & 0xFFFF0000u);
}

// ---- the per-thread sums: one matrix row, kFNPX consecutive columns -------------------------------------------------
template <int FM>
struct FastPass {
    static constexpr int SPAN = 2 * FM + kFNPX;
    static constexpr int SPAN_DN = FM + kFNPX - 1;     // the lower-left boxes only reach the columns left of the last pixel
    static constexpr int PX = fast_px(FM);

    // xrow = &tile[row of r][8 * column block]: element (r + a, c0 - FM + t) is xrow[a * PX + FM - a + t] (the tile is
    // indexed by diagonal along a row: one matrix row down is one diagonal back)
    template <int G>
    static __host__ __device__ __forceinline__ void vert(const float* __restrict__ xrow, float (&dn)[SPAN_DN], float (&W)[SPAN]) {
        {
            constexpr int c0 = FM - G, sh = c0 & 3, nv = (sh + SPAN + 3) / 4;
            float f[4 * nv];
            const float* p = xrow + G * PX + (c0 - sh);
#ifdef __CUDA_ARCH__
            sfor<0, nv>([&](auto I) {
                constexpr int i = decltype(I)::value;
                const float4 v = reinterpret_cast<const float4*>(p)[i];
                f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
            });
#else
            for (int i = 0; i < 4 * nv; ++i) f[i] = p[i];
#endif
            sfor<0, SPAN>([&](auto T) {
                constexpr int t = decltype(T)::value;
                if constexpr (G == 1) {
                    W[t] = f[sh + t];
                    if constexpr (t < SPAN_DN) dn[t] = f[sh + t];
                } else {
                    W[t] += f[sh + t];
                    if constexpr (t < SPAN_DN) dn[t] += f[sh + t];
                }
            });
        }
        {
            constexpr int c0 = FM + G, sh = c0 & 3, nv = (sh + SPAN + 3) / 4;
            float f[4 * nv];
            const float* p = xrow - G * PX + (c0 - sh);
#ifdef __CUDA_ARCH__
            sfor<0, nv>([&](auto I) {
                constexpr int i = decltype(I)::value;
                const float4 v = reinterpret_cast<const float4*>(p)[i];
                f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
            });
#else
            for (int i = 0; i < 4 * nv; ++i) f[i] = p[i];
#endif
            sfor<0, SPAN>([&](auto T) {
                constexpr int t = decltype(T)::value;
                W[t] += f[sh + t];
            });
        }
    }

    // quadrant boxes of half-width w for the 8 pixels: Q[i] = Q_w(r, c0 + i), L[i] = LL_w(r, c0 + i); fq / fl = the largest
    // full window sum / lower-left sum among them (bounds every intermediate of the slides)
    template <int w>
    static __host__ __device__ __forceinline__ void horiz(const float (&dn)[SPAN_DN], const float (&W)[SPAN], float (&Q)[kFNPX],
                                                          float (&L)[kFNPX], float& fq, float& fl) {
        float full = W[FM - w];
        sfor<FM - w + 1, FM + w + 1>([&](auto T) { full += W[decltype(T)::value]; });
        float ll = dn[FM - w];
        sfor<FM - w + 1, FM>([&](auto T) { ll += dn[decltype(T)::value]; });
        float mq = full, ml = ll;
        sfor<0, kFNPX>([&](auto I) {
            constexpr int i = decltype(I)::value;
            Q[i] = full - W[FM + i];
            L[i] = ll;
            if constexpr (i + 1 < kFNPX) {
                full = (full - W[FM - w + i]) + W[FM + w + 1 + i];
                ll = (ll - dn[FM - w + i]) + dn[FM + i];
                mq = fast_max(mq, full);
                ml = fast_max(ml, ll);
            }
        });
        fq = mq;
        fl = ml;
    }

    // lvpk: 8 nibbles, the level code (step index) of each pixel, 0xF = not a resolved pixel; lvmask: steps present in the
    // warp; ft: largest half-width any pixel of the warp needs.  For every resolved pixel i the sink receives, at the
    // pixel's own width w = W0 + code:  sink(i, code, K = Q_w - Q_p, Y = LL_w - LL_p, bounds on |K' - K| and |Y' - Y| as two bf16).
    // P, W0: the (p, w) pair of the run (run-time values: one kernel covers every single-pair program up to maxww = FM);
    // CP, CW >= 0: the pair is known at compile time (the usual pairs get their own kernels: only the window code they can
    // reach is generated -- the unrolled sums are ~60 KB of code and instruction-cache misses are a measurable stall)
    template <int CP, int CW, class Sink>
    static __host__ __device__ __forceinline__ void run(const float* __restrict__ xrow, int P, int W0, unsigned lvpk, unsigned lvmask,
                                                        int ft, Sink&& sink) {
        float dn[SPAN_DN], W[SPAN];
        float Qp[kFNPX], Lp[kFNPX], fqp = 0.f, flp = 0.f;
        sfor<0, kFNPX>([&](auto I) { Qp[decltype(I)::value] = 0.f; Lp[decltype(I)::value] = 0.f; });
        sfor<1, FM + 1>([&](auto GG) {
            constexpr int g = decltype(GG)::value;
            if (g <= ft) {
                vert<g>(xrow, dn, W);
                auto level = [&]() {
                    float Q[kFNPX], L[kFNPX], fq, fl;
                    horiz<g>(dn, W, Q, L, fq, fl);
                    // the bounds of this width, kept as two bf16 rounded up (one word of the record)
                    const unsigned epk = fast_pack_err(fast_cerr_k(g) * (fq + fqp), fast_cerr_y(g) * (fl + flp));
                    const unsigned sc = (unsigned)(g - W0);
                    sfor<0, kFNPX>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        if (((lvpk >> (4 * i)) & 0xFu) == sc) sink(I, sc, Q[i] - Qp[i], L[i] - Lp[i], epk);
                    });
                };
                if constexpr (CP >= 0) {
                    if constexpr (g == CP) horiz<g>(dn, W, Qp, Lp, fqp, flp);
                    else if constexpr (g >= CW) { if ((lvmask >> (g - CW)) & 1u) level(); }
                } else {
                    if (g <= kFMaxPeak && g == P) {     // (peak widths beyond kFMaxPeak take the general-form kernel)
                        if constexpr (g <= kFMaxPeak) horiz<g>(dn, W, Qp, Lp, fqp, flp);
                    } else if (g >= kFMinWidth && g >= W0 && ((lvmask >> (g - W0)) & 1u)) {
                        if constexpr (g >= kFMinWidth) level();
                    }
                }
            }
        });
    }
};

// The general form of FastPass::run: K = sum_g c[code][g] Q_g, Y = sum_g c[code][g] LL_g (coefficients in shared memory,
// looked up per pixel), window sums only at the rings some code present in the warp needs (hm).
template <int FM>
struct FastPassGen {
    using FP = FastPass<FM>;
    static constexpr int SPAN = FP::SPAN;
    template <class Sink>
    static __host__ __device__ __forceinline__ void run(const float* __restrict__ xrow, int W0, unsigned lvpk, unsigned lvmask, int ft,
                                                        unsigned hm, const float* __restrict__ ctab, const float* __restrict__ cabs,
                                                        Sink&& sink) {
        float dn[FP::SPAN_DN], W[SPAN];
        float Ka[kFNPX], Ya[kFNPX], ek = 0.f, ey = 0.f;
        sfor<0, kFNPX>([&](auto I) { Ka[decltype(I)::value] = 0.f; Ya[decltype(I)::value] = 0.f; });
        sfor<1, FM + 1>([&](auto GG) {
            constexpr int g = decltype(GG)::value;
            if (g <= ft) {
                FP::template vert<g>(xrow, dn, W);
                if ((hm >> g) & 1u) {
                    float Q[kFNPX], L[kFNPX], fq, fl;
                    FP::template horiz<g>(dn, W, Q, L, fq, fl);
                    ek += cabs[g] * fast_cerr_gen_k(g) * fq;
                    ey += cabs[g] * fast_cerr_gen_y(g) * fl;
                    const bool res = g >= W0 && ((lvmask >> (g - W0)) & 1u);
                    const unsigned epk = fast_pack_err(ek, ey);
                    sfor<0, kFNPX>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        const unsigned code = (lvpk >> (4 * i)) & 0xFu;
                        if (code != 0xFu) {
                            const float c = ctab[code * kFMaxG + g];
                            Ka[i] = fmaf(c, Q[i], Ka[i]);
                            Ya[i] = fmaf(c, L[i], Ya[i]);
                            if (res && code == (unsigned)(g - W0)) sink(I, code, Ka[i], Ya[i], epk);
                        }
                    });
                }
            }
        });
    }
};

// lambda-chunk edges inside one octave: chunk i has the upper edge 2^((i-1)/3), so for x in [2^e, 2^(e+1)) the chunk is
// 3e + 2 + (mantissa >= 2^(1/3)) + (mantissa >= 2^(2/3)) -- an exponent and two mantissa compares, no table.  c?dn / c?up:
// mantissa bits of the two edges rounded down / up (with a 1e-9 margin for the edges the caller's pow() produced;
// hp_ctx_create checks that every edge is 2^e times one of the two within 1e-12).
struct FastEdges { unsigned c1dn, c2dn, c1up, c2up; };
__host__ __device__ __forceinline__ unsigned fast_bits(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    union { float f; unsigned u; } v; v.f = x; return v.u;
#endif
}
__host__ __device__ __forceinline__ int fast_chunk_index(unsigned bits, unsigned c1, unsigned c2) {      // bits of a positive float
    const int e = (int)(bits >> 23) - 127;
    const unsigned m = bits & 0x7fffffu;
    const int i = 3 * e + 2 + (m >= c1 ? 1 : 0) + (m >= c2 ? 1 : 0);
    return i < 1 ? 1 : i;                               // below 1: chunk 1 = (0, 1)
}

// Classification of one background of one pixel from its fp32 sum S (|S - exact| <= es), the fp32 factor f = IR / bE
// and bb = B1 * B2.  0: certainly not a valid pixel (E == 0), 1: certainly valid and strictly inside lambda-chunk
// `chunk` (chunk == maxchunk + 1: beyond the last edge), 2: cannot tell -- evaluate exactly.
// Certain means: with i = the chunk of hi (edges rounded DOWN, so hi < edge(i) for sure), lo lies above the lower edge
// of chunk i rounded UP (cinfo[i].w, float bits; 0 for chunk 1): edge(i - 1) < lo <= E <= hi < edge(i) for every E the
// bound allows.  inf = cinfo[chunk] = {first histogram bin, bins, candidate threshold, lower edge}.
__host__ __device__ __forceinline__ int fast_classify(float S, float es, float f, float bb, const FastEdges& ed, int mc,
                                                      const int4* __restrict__ cinfo, int& chunk, float& lo, float& hi, int4& inf) {
    // straight-line on purpose (no early returns): the donut and the lower-left classification of a record are independent
    // chains and the compiler interleaves them only when neither sits behind a branch
    const bool zero = f == 0.f || bb == 0.f || es == 0.f;   // bE == 0 / IR == 0 / a zero bias, or every cell in reach is zero
    const float fm = f * bb;
    const float ea = S * fm;
    const float err = es * fabsf(fm) + fabsf(ea) * kFCRel;
    lo = ea - err;
    hi = ea + err;
    const int ih = fast_chunk_index(fast_bits(hi), ed.c1dn, ed.c2dn);      // (garbage in, some index out: clamped below)
    const int ic = ih > mc ? mc + 1 : ih;
    inf = cinfo[ic];
#ifdef __CUDA_ARCH__
    const float elo = __int_as_float(inf.w);
#else
    float elo; { union { int i; float f; } v; v.i = inf.w; elo = v.f; }
#endif
    const bool sure = hi < 1e37f && lo > elo;            // (false for lo <= 0 and for NaN)
    const int code = zero ? 0 : (sure ? 1 : 2);
    chunk = code == 1 ? ic : 0;
    if (zero) { lo = 0.f; hi = 0.f; }
    return code;
}

// shared-memory layout of a k_score_fast CTA (offsets in bytes): kFStages tile stages, then the per-CTA state
template <int FM, int NEX>
struct FastLayout {
    static constexpr size_t al(size_t x) { return (x + 127) & ~(size_t)127; }
    static constexpr int PX = fast_px(FM);
    // one stage (everything in it arrives by TMA / bulk copies on the stage's mbarrier)
    static constexpr size_t sXS = 0;                                             // float [kFXR][PX] fp32 balanced tile
    static constexpr size_t sOBS = al(sXS + (size_t)kFXR * PX * 4);              // int   [kFTD][4][kFTR / 4] raw counts
    static constexpr size_t sLVL = al(sOBS + (size_t)kFTD * kFTR * 4);           // u8    [kFTD][4][kFTR / 4] levels
    static constexpr size_t sFTAB = al(sLVL + (size_t)kFTD * kFTR);              // float [2][nexec][kFTD] (nexec <= NEX)
    static constexpr size_t sB1 = al(sFTAB + (size_t)2 * NEX * kFTD * 4);        // float [kFTR]
    static constexpr size_t sB2 = al(sB1 + kFTR * 4);                            // float [kFTR + kFTD]
    static constexpr size_t stage = al(sB2 + (kFTR + kFTD) * 4);
    static constexpr size_t oHIST = kFStages * stage;                            // u32   [2][kShI][kShK]
    static constexpr size_t oQ = al(oHIST + (size_t)2 * kShI * kShK * 4);        // uint4 [warps][kFQCap] K, Y, err pack, meta
    static constexpr size_t oCINFO = al(oQ + (size_t)kFWarps * kFQCap * 16);     // int4  [kChunkTab]
    static constexpr size_t oMISC = al(oCINFO + (size_t)kChunkTab * 16);         // runmax[2], counters, mbarriers
    static constexpr size_t oGEN = oMISC + 128;                                  // general form: float ctab[16][16], cabs[16]; u8 tcode, tstep
    static constexpr size_t bytes = oGEN + (size_t)kFMaxCode * kFMaxG * 4 + kFMaxG * 4 + ((HP_MAX_STEPS + 2 + 127) & ~127) + kFMaxCode;
    static constexpr uint32_t tx_fixed = (uint32_t)(kFXR * PX * 4 + kFTD * kFTR * 4 + kFTD * kFTR + kFTR * 4 + (kFTR + kFTD) * 4);
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_ld_volatile(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void smem_st_volatile(unsigned* p, unsigned v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void gmem_red_add(unsigned* p, unsigned v) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void gmem_red_max(unsigned* p, unsigned v) {
    asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void smem_red_max32(unsigned* p, unsigned v) {
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completing on an mbarrier; 16-byte aligned addresses and size
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int FM, bool GEN, int CP = -1, int CW = -1>
__global__ void __launch_bounds__(kFThreads, 1) k_score_fast(const __grid_constant__ CUtensorMap tm_raw, const __grid_constant__ CUtensorMap tm_x,
                                                             const __grid_constant__ CUtensorMap tm_lvl, const __grid_constant__ FastArgs A) {
    using FP = FastPass<FM>;
    constexpr int PX = FP::PX;
    constexpr int NEX = FM;                            // widths a pair resolves at: at most FM - w + 1
    const int P = CP >= 0 ? CP : A.p, W0 = CW >= 0 ? CW : A.w0, ncode = A.ncode;
    using LY = FastLayout<FM, NEX>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned int* const shist = reinterpret_cast<unsigned int*>(smem + LY::oHIST);
    int4* const cinfo = reinterpret_cast<int4*>(smem + LY::oCINFO);
    unsigned int* const runmax = reinterpret_cast<unsigned int*>(smem + LY::oMISC);          // [2]
    unsigned int* const nscr = reinterpret_cast<unsigned int*>(smem + LY::oMISC + 8);
    unsigned int* const claim = reinterpret_cast<unsigned int*>(smem + LY::oMISC + 12);
    unsigned int* const done = reinterpret_cast<unsigned int*>(smem + LY::oMISC + 16);       // [kFStages]
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(smem + LY::oMISC + 32);           // [kFStages]
    float* const s_ctab = reinterpret_cast<float*>(smem + LY::oGEN);
    float* const s_cabs = s_ctab + kFMaxCode * kFMaxG;
    unsigned char* const s_tcode = reinterpret_cast<unsigned char*>(s_cabs + kFMaxG);
    unsigned char* const s_tstep = s_tcode + ((HP_MAX_STEPS + 2 + 127) & ~127);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    const int nexec = A.nexec, n = A.n, num = A.num, mc = A.maxchunk;
    uint4* const q = reinterpret_cast<uint4*>(smem + LY::oQ) + warp * kFQCap;
    const FastEdges ed{A.c1dn, A.c2dn, A.c1up, A.c2up};
    // this CTA's tiles: items blockIdx.x, blockIdx.x + gridDim.x, ...  Item order: the strip next to the diagonal first
    // (E.max() lives there: once its lower bound is known, no pixel of the other strips is sent to the exact list as a
    // contender), then the far strips (the deepest levels, the longest tiles) first.
    const int items = A.nstrips * A.ntr;
    const int ntiles = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_of = [&](int j, int& strip, int& r0) {
        const int item = (int)blockIdx.x + j * (int)gridDim.x;
        const int sk = item / A.ntr;
        strip = sk == 0 ? 0 : A.nstrips - sk;
        r0 = (item - sk * A.ntr) * kFTR;
    };
    // one thread: every copy of tile j onto the mbarrier of stage j & 1
    auto issue = [&](int j) {
        int strip, r0;
        tile_of(j, strip, r0);
        unsigned char* const sb = smem + (size_t)(j & 1) * LY::stage;
        uint64_t* const bar = full_bar + (j & 1);
        const uint32_t ftb = (uint32_t)(2 * ncode * kFTD * 4);
        mbar_expect_tx(bar, LY::tx_fixed + ftb);
        tma_load_2d(sb + LY::sXS, &tm_x, strip * kFTD - 2 * FM, r0 - kFRowHalo, bar);
        tma_load_3d(sb + LY::sOBS, &tm_raw, r0 / 4, 0, A.dlo + strip * kFTD, bar);
        tma_load_3d(sb + LY::sLVL, &tm_lvl, r0 / 4, 0, A.dlo + strip * kFTD, bar);
        bulk_load(sb + LY::sFTAB, A.ffs + (size_t)strip * 2 * ncode * kFTD, ftb, bar);
        bulk_load(sb + LY::sB1, A.b1f + r0, kFTR * 4, bar);
        bulk_load(sb + LY::sB2, A.b2s + r0 + strip * kFTD, (kFTR + kFTD) * 4, bar);
    };
    {
        const Chunks& C = A.tab->chunks;
        for (int i = tid; i < kChunkTab; i += kFThreads) {
            const bool in = i <= C.maxchunk;
            const int eb = i <= C.maxchunk + 1 ? __float_as_int(A.elo[i]) : 0x7f800000;     // lower edge of chunk i, rounded up
            cinfo[i] = in ? make_int4(C.hoff[i], C.hw[i], C.kcand[i], eb) : make_int4(0, 1, 0x7fffffff, eb);
        }
        for (int i = tid; i < 2 * kShI * kShK; i += kFThreads) shist[i] = 0u;
        if constexpr (GEN) {
            for (int i = tid; i < kFMaxCode * kFMaxG; i += kFThreads) s_ctab[i] = A.ctab[i / kFMaxG][i % kFMaxG];
            for (int i = tid; i < kFMaxG; i += kFThreads) s_cabs[i] = A.cabs[i];
            for (int i = tid; i < HP_MAX_STEPS + 2; i += kFThreads) s_tcode[i] = A.tcode[i];
        }
        for (int i = tid; i < kFMaxCode; i += kFThreads) s_tstep[i] = A.tstep[i];
        if (tid == 0) {
            runmax[0] = 0u; runmax[1] = 0u; *nscr = 0u; *claim = 0u;
            for (int s = 0; s < kFStages; ++s) { done[s] = 0u; mbar_init(full_bar + s, 1); }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int j = 0; j < kFStages && j < ntiles; ++j) issue(j);
        }
    }
    __syncthreads();
    unsigned nvk = 0, nvy = 0;                          // valid pixels this thread classified (K / Y)
    unsigned cbase = 0, cused = kFCandChunk;            // this warp's piece of the candidate list (warp-uniform)

#pragma unroll 1
    for (;;) {
        unsigned g = 0;
        if (lane == 0) g = smem_atom_add(claim, 1u);
        g = __shfl_sync(full, g, 0);
        const int j = (int)(g / kFPasses), ps = (int)(g % kFPasses);
        if (j >= ntiles) break;
        const int st = j & 1;
        int strip, r0;
        tile_of(j, strip, r0);
        const int d0 = A.dlo + strip * kFTD;
        const bool edge_tile = r0 < A.F || r0 + d0 + kFTR + kFTD - 2 >= n - A.F;      // some pixel of the tile is next to a chromosome end
        unsigned char* const sb = smem + (size_t)st * LY::stage;
        const float* const xs = reinterpret_cast<const float*>(sb + LY::sXS);
        const int* const obs = reinterpret_cast<const int*>(sb + LY::sOBS);
        const unsigned char* const lvt = sb + LY::sLVL;
        const float* const ftab = reinterpret_cast<const float*>(sb + LY::sFTAB);
        const float* const b1t = reinterpret_cast<const float*>(sb + LY::sB1);
        const float* const b2t = reinterpret_cast<const float*>(sb + LY::sB2);
        mbar_wait(full_bar + st, (unsigned)(j >> 1) & 1u);

        // ---- the sums of one pass: 32 rows (one per lane) x 8 columns ---------------------------------------------
        const int cb = ps >> 1;
        const int rl = 32 * (ps & 1) + lane;           // this thread's row of the tile
        unsigned lvpk = 0, mine = 0, slotpk0 = 0, slotpk1 = 0;
        int cnt = 0;                                    // records of this warp and pass (warp-uniform)
        {
            const unsigned char* lp = lvt + (kFNPX * cb * 4 + (rl & 3)) * (kFTR / 4) + (rl >> 2);
            const int r = r0 + rl;
#pragma unroll
            for (int i = 0; i < kFNPX; ++i) {
                // stale bytes beyond the chromosome or the band never reach here as levels: (r, d) is checked
                const int d = d0 + kFNPX * cb + i;
                unsigned lv = lp[i * kFTR];
                if constexpr (GEN) lv = (lv < (unsigned)A.nlevels && d >= A.wpair) ? (unsigned)s_tcode[lv] : 0xFu;
                const unsigned code = (lv < (unsigned)ncode && d <= A.dhi && r < n && r + d < n) ? lv : 0xFu;
                lvpk |= code << (4 * i);
                const bool e = code != 0xFu;
                if (e) mine |= 1u << code;
                const unsigned mb = __ballot_sync(full, e);
                const unsigned slot = (unsigned)cnt + __popc(mb & lt);       // < 256
                if (i < 4) slotpk0 |= slot << (8 * i); else slotpk1 |= slot << (8 * (i - 4));
                cnt += __popc(mb);
            }
        }
        const unsigned lvmask = __reduce_or_sync(full, mine);
        if (lvmask != 0u) {                             // (else: no resolved pixel in these 32 rows x 8 columns)
            const int ft = W0 + (31 - __clz((int)lvmask));
            // element (r + a, c0 - FM + t) sits at tile column (d0 + 8 cb - FM + t - a) - (d0 - 2 FM) = 8 cb + FM + t - a
            const float* xrow = xs + (size_t)(rl + kFRowHalo) * PX + kFNPX * cb;
            const unsigned mbase = (unsigned)rl | ((unsigned)(kFNPX * cb) << 6);
            auto push = [&](auto I, unsigned sc, float kv, float yv, unsigned epk) {
                constexpr int i = decltype(I)::value;
                const unsigned slot = ((i < 4 ? slotpk0 : slotpk1) >> (8 * (i & 3))) & 0xffu;
                q[slot] = make_uint4(__float_as_uint(kv), __float_as_uint(yv), epk, mbase + ((unsigned)i << 6) + (sc << 12));
            };
            if constexpr (GEN) {
                unsigned hm = 0;                        // rings at which some code present in the warp has a coefficient
                for (unsigned mm = lvmask; mm; mm &= mm - 1u) hm |= A.hmask[__ffs((int)mm) - 1];
                FastPassGen<FM>::run(xrow, W0, lvpk, lvmask, ft, hm, s_ctab, s_cabs, push);
            } else {
                FP::template run<CP, CW>(xrow, P, W0, lvpk, lvmask, ft, push);
            }
            __syncwarp();
            // ---- close the records: classify both backgrounds, account the certain ones ---------------------------
#pragma unroll 1
            for (int b0 = 0; b0 < cnt; b0 += 32) {
                const bool act = b0 + lane < cnt;
                const uint4 rec = act ? q[b0 + lane] : make_uint4(0u, 0u, 0u, 0u);
                const unsigned m = rec.w;
                const int rrl = m & 63, dl = (m >> 6) & 63, sc = (m >> 12) & 15;
                const int s = s_tstep[sc];              // the executed step the record resolves at
                const int r = r0 + rrl, d = d0 + dl;
                const int ob = obs[(dl * 4 + (rrl & 3)) * (kFTR / 4) + (rrl >> 2)];
                float fk = ftab[sc * kFTD + dl], fy = ftab[(ncode + sc) * kFTD + dl];
                const bool top = edge_tile && r < A.F, end = edge_tile && r + d >= n - A.F;
                if (act && (top || end)) {              // next to a chromosome end: the factor tables of that row / column
                    const int z = top ? 1 + r : 1 + A.F + (n - 1 - r - d);
                    const size_t at = ((size_t)(z * 2) * nexec + s) * num + d;
                    fk = (top && end) ? NAN : A.ffac[at];                    // next to both: bE is a walk over the cell list
                    fy = (top && end) ? NAN : A.ffac[at + (size_t)nexec * num];
                }
                const float bb = b1t[rrl] * b2t[rrl + dl];
                int ck, cy;
                float lo0, hi0, lo1, hi1;
                int4 infk, infy;
                const int c0 = fast_classify(__uint_as_float(rec.x), __uint_as_float(rec.z << 16), fk, bb, ed, mc, cinfo, ck, lo0, hi0, infk);
                const int c1 = fast_classify(__uint_as_float(rec.y), __uint_as_float(rec.z & 0xFFFF0000u), fy, bb, ed, mc, cinfo, cy, lo1, hi1, infy);
                const bool ex = act && (c0 == 2 || c1 == 2 || !(bb == bb));
                const bool ok = act && !ex;
                unsigned flags = 0, emk = 0;
                bool cand = false;
                if (ok && c0 == 1) {
                    flags |= HP_SF_VALID_K;
                    ++nvk;
                    if (ck > mc) {
                        gmem_red_add(&A.cnt[2], 1u);
                    } else {
                        const int4 inf = infk;
                        const int kb = ob < inf.y - 1 ? ob : inf.y - 1;
                        if (ck <= kShI && kb < kShK) smem_red_add(&shist[(ck - 1) * kShK + kb], 1u);
                        else gmem_red_add(&A.hist[(size_t)inf.x + kb], 1u);
                        cand |= ob >= inf.z;
                    }
                    // E.max(): a record whose interval reaches the largest lower bound seen so far is evaluated exactly
                    const unsigned cur = smem_ld_volatile(&runmax[0]);
                    if (__float_as_uint(hi0) >= cur) emk |= 1u;
                    if (__float_as_uint(lo0) > cur) smem_red_max32(&runmax[0], __float_as_uint(lo0));
                }
                if (ok && c1 == 1) {
                    flags |= HP_SF_VALID_Y | HP_SF_CEMY_NONZERO;
                    ++nvy;
                    if (cy > mc) {
                        gmem_red_add(&A.cnt[2], 1u);
                    } else {
                        const int4 inf = infy;
                        const int kb = ob < inf.y - 1 ? ob : inf.y - 1;
                        if (cy <= kShI && kb < kShK) smem_red_add(&shist[(kShI + cy - 1) * kShK + kb], 1u);
                        else gmem_red_add(&A.hist[(size_t)A.total_bins + inf.x + kb], 1u);
                        cand |= ob >= inf.z;
                    }
                    const unsigned cur = smem_ld_volatile(&runmax[1]);
                    if (__float_as_uint(hi1) >= cur) emk |= 2u;
                    if (__float_as_uint(lo1) > cur) smem_red_max32(&runmax[1], __float_as_uint(lo1));
                }
                // candidates (classified records whose Poisson tail can pass sig) go to the warp's own piece of the
                // list: one global atomic per kFCandChunk records instead of one (with its round trip) per batch
                const unsigned mcand = __ballot_sync(full, cand);
                if (mcand) {
                    const unsigned nc = __popc(mcand);
                    if (cused + nc > kFCandChunk) {
                        if (cused < kFCandChunk && lane < kFCandChunk - cused && cbase + cused + lane < A.fcand_cap)
                            A.fcand[cbase + cused + lane].r = -1;            // (a warp leaves at most 31 slots behind)
                        if (cused < kFCandChunk && kFCandChunk - cused > 32 && lane + 32 < kFCandChunk - cused &&
                            cbase + cused + lane + 32 < A.fcand_cap)
                            A.fcand[cbase + cused + lane + 32].r = -1;
                        unsigned nb = 0;
                        if (lane == 0) nb = atomicAdd(&A.cnt[8], kFCandChunk);
                        cbase = __shfl_sync(full, nb, 0);
                        cused = 0;
                    }
                    if (cand) {
                        const unsigned slot = cbase + cused + __popc(mcand & lt);
                        if (slot < A.fcand_cap)
                            *reinterpret_cast<int4*>(&A.fcand[slot]) =
                                make_int4(r, d | (s << 16), ob, (int)((unsigned)ck | ((unsigned)cy << 8) | (flags << 16) | ((unsigned)A.pair << 24)));
                        else gmem_red_add(&A.cnt[9], 1u);
                    }
                    cused += nc;
                }
                // records the exact kernel must evaluate and account
                const unsigned mx = __ballot_sync(full, ex);
                if (mx) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&A.cnt[10], (unsigned)__popc(mx));
                    base = __shfl_sync(full, base, 0);
                    if (ex) {
                        const unsigned slot = base + __popc(mx & lt);
                        if (slot < A.xrec_cap) *reinterpret_cast<int4*>(&A.xrec[slot]) = make_int4(r, d | (s << 16), ob, 0);
                        else gmem_red_add(&A.cnt[11], 1u);
                    }
                }
                // E.max() contenders wait in the CTA's scratch list: most of them fall below the bound the CTA ends with
                const unsigned me = __ballot_sync(full, emk != 0u);
                if (me) {
                    unsigned base = 0;
                    if (lane == 0) base = smem_atom_add(nscr, (unsigned)__popc(me));
                    base = __shfl_sync(full, base, 0);
                    if (emk) {
                        const unsigned slot = base + __popc(me & lt);
                        const int hk = (emk & 1u) ? __float_as_int(hi0) : 0, hy = (emk & 2u) ? __float_as_int(hi1) : 0;
                        if (slot < kFScratch) {
                            A.scratch[(size_t)blockIdx.x * kFScratch + slot] = make_int4(r, d | (s << 16), hk, hy);
                        } else {                       // scratch full: straight to the exact list
                            const unsigned gx = atomicAdd(&A.cnt[10], 1u);
                            if (gx < A.xrec_cap) *reinterpret_cast<int4*>(&A.xrec[gx]) = make_int4(r, d | (s << 16), ob, (int)emk);
                            else atomicAdd(&A.cnt[11], 1u);
                        }
                    }
                }
            }
        }
        // ---- this warp is done with its pass; the warp that finishes the last pass of the tile owns the stage: it
        // exchanges the running E.max() bounds with the other CTAs and re-arms the stage with the tile after next --------
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const unsigned old = smem_atom_add(done + st, 1u);
            if (old == (unsigned)(kFPasses - 1)) {
                __threadfence_block();
                smem_st_volatile(done + st, 0u);
                if (j + kFStages < ntiles) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(j + kFStages);
                }
                const unsigned g0 = A.gmax[0], g1 = A.gmax[1];          // (any value read is a valid bound)
                const unsigned l0 = smem_ld_volatile(&runmax[0]), l1 = smem_ld_volatile(&runmax[1]);
                if (g0 > l0) smem_red_max32(&runmax[0], g0); else if (l0 > g0) gmem_red_max(&A.gmax[0], l0);
                if (g1 > l1) smem_red_max32(&runmax[1], g1); else if (l1 > g1) gmem_red_max(&A.gmax[1], l1);
            }
        }
    }
    // ---- E.max() contenders that still reach the largest lower bound known now -> exact list ---------------------------
    __syncthreads();
    if (tid == 0) {
        if (runmax[0]) atomicMax(&A.gmax[0], runmax[0]);
        if (runmax[1]) atomicMax(&A.gmax[1], runmax[1]);
    }
    {
        const unsigned m0 = max(runmax[0], A.gmax[0]), m1 = max(runmax[1], A.gmax[1]);
        const unsigned ns = min(*nscr, kFScratch);
        for (unsigned i = tid; i < ns; i += kFThreads) {
            const int4 c = A.scratch[(size_t)blockIdx.x * kFScratch + i];
            const int kind = ((c.z != 0 && (unsigned)c.z >= m0) ? 1 : 0) | ((c.w != 0 && (unsigned)c.w >= m1) ? 2 : 0);
            if (kind) {
                const unsigned gx = atomicAdd(&A.cnt[10], 1u);
                if (gx < A.xrec_cap) *reinterpret_cast<int4*>(&A.xrec[gx]) = make_int4(c.x, c.y, 0, kind);
                else atomicAdd(&A.cnt[11], 1u);
            }
        }
    }
    // ---- flush: the unused rest of the warp's candidate piece, valid counts, privatised histogram ----------------------
    for (unsigned k = cused + lane; k < kFCandChunk; k += 32)
        if (cbase + k < A.fcand_cap) A.fcand[cbase + k].r = -1;
    {
        const unsigned a = __reduce_add_sync(full, nvk), b = __reduce_add_sync(full, nvy);
        if (lane == 0 && a) atomicAdd(&A.nvalid[0], (unsigned long long)a);
        if (lane == 0 && b) atomicAdd(&A.nvalid[1], (unsigned long long)b);
    }
    for (int i = tid; i < 2 * kShI * kShK; i += kFThreads) {
        const unsigned v = shist[i];
        if (v) {
            const int kb = i % kShK, ci = (i / kShK) % kShI + 1, fl = i / (kShK * kShI);
            atomicAdd(&A.hist[(size_t)fl * A.total_bins + cinfo[ci].x + kb], v);
        }
    }
}

// ---- upload-time companions: the fp32 row-major plane and the fp32 bias vectors the kernel's TMA / bulk copies read ----
// xf[r][d - bf] = (float)balanced(r, r + d), row pitch ndp floats (a multiple of 4: 16-byte rows for the tensor map);
// columns beyond the band and elements beyond the chromosome are zero (the fp64 plane has zero tails).
__global__ void __launch_bounds__(256) k_f32plane(const double* __restrict__ bal, float* __restrict__ xf, int n, int num, int bf, int pitch,
                                                  int ndp) {
    __shared__ float t[32][65];                         // [diagonal][row]
    const int r0 = blockIdx.x * 64, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int dd = ty; dd < 32; dd += 8) {
        const int d = bf + k0 + dd;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + 32 * h + tx;             // 32 consecutive rows of one plane: four 64-byte pieces
            t[dd][32 * h + tx] = (d < num && r < pitch) ? (float)bal[qidx(d, r, pitch)] : 0.f;
        }
    }
    __syncthreads();
    for (int rr = ty; rr < 64; rr += 8) {
        const int r = r0 + rr, k = k0 + tx;
        if (r < n && k < ndp) xf[(size_t)r * ndp + k] = t[tx][rr];
    }
}
__global__ void k_fast_bias(const double* __restrict__ b1, const double* __restrict__ b2, float* __restrict__ b1f, float* __restrict__ b2s,
                            int n, int bf, int n1, int n2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) b1f[i] = i < n ? fast_bias(b1[i]) : 0.f;
    if (i < n2) b2s[i] = i + bf < n ? fast_bias(b2[i + bf]) : 0.f;
}
#endif  // __CUDACC__

}  // namespace hp
