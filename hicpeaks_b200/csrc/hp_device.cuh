// Device-side tables, small PTX helpers (mbarrier + TMA) and shared structs of the HiCCUPS engine.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/hicpeaks_b200.h"

namespace hp {

constexpr int kTR = 128;          // tile rows (matrix row index r)
constexpr int kThreads = 256;
constexpr int kHR = 16;           // row halo of a balanced tile: >= maxww of the specialised kernels and a multiple of 8
                                  // (a TMA box must start on a 16-byte boundary: an even row quad)
constexpr int kNQ = (kTR + 2 * kHR) / 4;   // row quads of a balanced tile (40)
constexpr int kNQL = (kTR + 16) / 4;       // row quads of a raw tile of the specialised level kernel (rows r .. r + maxww + 3)
constexpr int kMaxOps = 6144;     // stencil offsets of one whole sweep program
constexpr int kMaxROps = 512;     // offsets that feed Reads (<= maxww^2)
constexpr int kMaxChunk = 64;     // lambda-chunk cap
constexpr int kShI = 20;          // smem-privatised histogram: chunks 1..kShI
constexpr int kShK = 64;          //                            bins 0..kShK-1
constexpr int kShPairs = 4;
constexpr int kStage = 512;       // per-CTA candidate staging records
constexpr unsigned char kLvlNever = 0xFE, kLvlNone = 0xFF, kNoStep = 0xFF;

struct Prog {                     // the sweep program (callers.py:132-198 unrolled by the host)
    int nsteps, nsteps_exec, npw, thr;
    int pw[HP_MAX_PW], ww[HP_MAX_PW];
    int step_pi[HP_MAX_STEPS], step_w[HP_MAX_STEPS];
    int op_end[HP_MAX_STEPS];     // ops [op_end[s-1], op_end[s]) belong to step s
    int rop_end[HP_MAX_STEPS];
    unsigned char next_step[HP_MAX_PW][HP_MAX_STEPS + 2];  // [pair][s*] -> executed step resolving it
    // [min(d - min(ww), dspan)][s*] -> last executed step a pixel on diagonal d with level s* needs (kNoStep: none)
    unsigned char last_need[HP_MAX_WW + 1][HP_MAX_STEPS + 2];
    int dspan;                    // max(ww) - min(ww)
    unsigned char step_lo[HP_MAX_STEPS];   // executed step s resolves its pair for levels in [step_lo[s], s]
};

struct Chunks {                   // lambda-chunk geometry (callers.py:30-38) + table layout
    double rv[kMaxChunk + 2];     // rv[i] = upper edge of chunk i (1-based); rv[0] = 0
    int hoff[kMaxChunk + 2];      // first bin of chunk i in the flat tables
    int hw[kMaxChunk + 2];        // bins of chunk i (observed counts >= hw-1 share the last bin, p == 0)
    int kcand[kMaxChunk + 2];     // smallest observed count with p <= sig (candidate threshold)
    int maxchunk, total_bins;
};

struct Tables {                   // per-context tables in global memory (one H2D copy per upload)
    Prog prog;
    Chunks chunks;
    signed char opa[kMaxOps];     // row / column offset of every cell the sweep adds, in fp64 addition order
    signed char opb[kMaxOps];
    unsigned char opy[kMaxOps];   // the cell also feeds the lower-left (Y) sums
    unsigned short ropi[kMaxROps];  // indices of the cells that feed Reads
};

struct Cand {                     // 32 B: a pixel whose Poisson p can pass sig for K or Y
    int r, d, obs;
    unsigned char pair, flags, chunk_k, chunk_y;
    double e_k, e_y;
};

// Band planes in HBM (raw counts, balanced values, levels): "quad-interleaved" diagonal planes.
// Element (r, r + d) lives at [d][r & 3][r >> 2] (sub-plane pitch = pitch / 4): a thread that owns four
// consecutive matrix rows then reads consecutive shared-memory words across the lanes of a warp for
// every row offset, and a 3-D TMA box (row quads, 4, diagonals) lands a tile in that order.
__host__ __device__ __forceinline__ size_t qidx(int d, int r, int pitch) {
    return (size_t)d * pitch + (size_t)(r & 3) * (pitch >> 2) + (r >> 2);
}

// fp32 factors of the re-associated score kernel (hp_score_fast.cuh): 0 = the pixel is certainly not valid
// (bE == 0, IR == 0 / NaN, bias 0 / NaN), NaN = fp32 cannot hold the factor safely (the record is evaluated exactly)
__host__ __device__ __forceinline__ float fast_factor(double ird, double be) {
    if (!(be > 0.0) && !(be < 0.0)) return 0.f;          // bE == 0 (or NaN: E is NaN, never valid)
    if (!(ird > 0.0) && !(ird < 0.0)) return 0.f;
    const double q = ird / be;
    const double aq = q < 0 ? -q : q;
    if (!(aq >= 1e-30 && aq <= 1e30)) return NAN;
    return (float)q;
}
__host__ __device__ __forceinline__ float fast_bias(double b) {
    if (!(b > 0.0) && !(b < 0.0)) return 0.f;            // 0 or NaN: E is 0 / NaN, never valid
    const double ab = b < 0 ? -b : b;
    if (!(ab >= 1e-15 && ab <= 1e15)) return NAN;
    return (float)b;
}

// Domain of the re-associated kernel's error bound: a balanced value is zero or a positive normal number in
// [2^-100, 2^100) (no sign, no NaN / Inf, nothing fp32 would flush).  Checked once per upload (k_relayout / k_prep_band).
__device__ __forceinline__ bool fast_domain_bad(double v) {
    const unsigned hw = (unsigned)__double2hiint(v), lw = (unsigned)__double2loint(v);
    return (hw - 0x39B00000u) >= 0x0C800000u && (hw | lw) != 0u;
}

// ---- mbarrier / TMA -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory atomics through explicit shared-space addresses: on a generic pointer the compiler emits the
// "which address space is this" fallback around every atomic (dozens of instructions and two branches each)
__device__ __forceinline__ unsigned smem_atom_add(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void smem_red_add(unsigned* p, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void smem_red_max(unsigned long long* p, unsigned long long v) {
    asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// 2-D tiled bulk tensor load global -> shared; out-of-bounds elements arrive as zeros
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

}  // namespace hp
