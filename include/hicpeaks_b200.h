/*
 * hicpeaks_b200 -- C ABI of the B200-native HiCCUPS scoring engine.
 *
 * This is the drop-in boundary for the hot path of XiaoTaoWang/HiCPeaks.  The reference has no FFI
 * (it is pure Python); the interface being replaced is the Python operator
 *
 *     hicpeaks.callers.hiccups(M, cM, B1, B2, IR, chromLen, Diags, cDiags, num, chrom, pw, ww, maxww,
 *                              sig, sumq, double_fold, single_fold, maxapart, res, use_raw,
 *                              min_marginal_peaks, onlyanchor, min_local_reads)
 *                                                      -- /root/reference/hicpeaks/callers.py:44-46
 *     hicpeaks.callers.bhfdr(...)                      -- callers.py:364-365
 *     hicpeaks.apa.apa_submatrix / apa_analysis        -- /root/reference/hicpeaks/apa.py:11,30
 *
 * as called by the per-chromosome worker (/root/reference/scripts/pyHICCUPS:139-175).  Each entry
 * point below cites the reference lines it replaces.  `hicpeaks_b200/callers.py` is the ctypes
 * binding that re-exposes the same Python signatures on top of this ABI.
 *
 * Conventions: plain C, no exceptions.  Every function returns HP_OK (0) or a negative hp_status;
 * hp_last_error(ctx) returns a message for the last failure on that context (ctx == NULL: the last
 * context-less failure of the calling thread).  Host buffers are caller-owned and only read during
 * the call; device memory is owned by the context.  A context is bound to one GPU and one CUDA
 * stream; it is not thread-safe, distinct contexts are independent.  There is NO CPU fallback: a
 * missing GPU / driver is an error.
 */
#ifndef HICPEAKS_B200_H
#define HICPEAKS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HP_ABI_VERSION 2
#define HP_MAX_PW 8        /* (pw, ww) pairs per run                                   */
#define HP_MAX_WW 20       /* largest supported maxww (reference function default: 20) */
#define HP_MAX_STEPS 160   /* sweep steps: sum over pairs of (maxww - ww + 1)          */

typedef enum hp_status {
    HP_OK = 0,
    HP_ERR_INVALID = -1,        /* bad argument                                                     */
    HP_ERR_CUDA = -2,           /* CUDA runtime / driver failure (message has the CUDA error)        */
    HP_ERR_NO_DEVICE = -3,      /* no usable sm_100 GPU                                              */
    HP_ERR_STATE = -4,          /* call sequence violated (e.g. score before upload)                 */
    HP_ERR_EMPTY_REFIDX = -5,   /* the reference raises here: unresolved set of some p became empty
                                   before the sweep ended, or no pixel at all (callers.py:205-208)    */
    HP_ERR_CHUNK_OVERFLOW = -6, /* an expected value needs a lambda-chunk beyond max_chunks           */
    HP_ERR_CAPACITY = -7        /* candidate / survivor buffer too small; retry with larger capacity  */
} hp_status;

typedef struct hp_ctx hp_ctx;

/* ---- library / context --------------------------------------------------------------------- */
int hp_abi_version(void);
int hp_device_count(int* count);
/* max_chunks: largest lambda-chunk index for which Poisson tables are kept (0 = default 52,
 * i.e. expected values up to 2^17; at most 64).  edges: NULL, or max_chunks doubles with the upper
 * edge rv_i of chunk i = 1..max_chunks exactly as the caller's numpy evaluates
 * np.power(2, (i - 1) / 3.) (callers.py:36-37); NULL uses the C library's pow(). */
int hp_ctx_create(int device, int max_chunks, const double* edges, hp_ctx** out);
void hp_ctx_destroy(hp_ctx* ctx);
const char* hp_last_error(const hp_ctx* ctx);

/* ---- input: the diagonal band of one chromosome --------------------------------------------- */
/* Replaces the containers the worker builds at scripts/pyHICCUPS:146-166 (Diags, M, cDiags, cM, IR,
 * biases).  Diagonal-major, exactly the reference's lists: raw_diags[d] has n - d int32 counts
 * (d in [0, num)); bal_diags[i] is balanced diagonal (bal_first + i) with n - bal_first - i doubles,
 * NaN already replaced by 0 (pyHICCUPS:157); ir[i] = IR[bal_first + i]; b1/b2 = biases (length n). */
typedef struct hp_band_desc {
    int64_t n;                        /* chromLen in bins                                   */
    int32_t num;                      /* stored raw diagonals, offsets 0 .. num-1           */
    int32_t bal_first;                /* first balanced offset == min(ww)                   */
    const int32_t* const* raw_diags;  /* [num]                                              */
    const double* const* bal_diags;   /* [num - bal_first]                                  */
    const double* ir;                 /* [num - bal_first]                                  */
    const double* b1;                 /* [n]                                                */
    const double* b2;                 /* [n]                                                */
} hp_band_desc;
int hp_band_upload(hp_ctx* ctx, const hp_band_desc* band);

/* Worker-level input (replaces the whole input preparation of scripts/pyHICCUPS:143-166): raw count
 * diagonals and the balancing weight of every bin (cooler's `bins[weight]`, NaN for masked bins).  The
 * balanced band ((w[r] * w[c]) * count as cooler evaluates bias1[row] * bias2[col] * data, NaN -> 0), IR[d] (mean of the non-NaN entries of balanced diagonal d,
 * bit-identical to numpy's mean) and the biases (1 / w, 0 where w is 0 or NaN) are computed on the device:
 * 4 bytes per band pixel cross PCIe instead of 12. */
typedef struct hp_counts_desc {
    int64_t n;
    int32_t num;
    int32_t bal_first;                /* min(ww)                                            */
    const int32_t* const* raw_diags;  /* [num]                                              */
    const double* weights;            /* [n]                                                */
} hp_counts_desc;
int hp_band_upload_counts(hp_ctx* ctx, const hp_counts_desc* band);
/* inspection of the uploaded / derived band: what = 0 IR[num], 1 biases[n], 2 balanced band [num][n] */
int hp_dump_band(hp_ctx* ctx, int32_t what, double* out, int64_t capacity);

/* Bytes the last hp_band_upload / hp_band_upload_counts moved host -> device.  hp_band_upload_counts narrows each
 * count diagonal on the host to the smallest of u8 / u16 / i32 that holds its values exactly before it crosses
 * PCIe (the device widens it again), so this is usually ~1 byte per band pixel.  Measurement aid (bench.py). */
int hp_upload_bytes(hp_ctx* ctx, int64_t* bytes);

/* Host-only inspection hook (needs neither a GPU nor a context): writes src[0..len) to dst in the narrowest of u8 / u16 /
 * i32 that holds every value exactly -- what hp_band_upload_counts does to each count diagonal before it crosses PCIe --
 * and returns the element size chosen (1, 2 or 4) in *esize.  dst must have room for len * 4 bytes. */
int hp_narrow_diagonal(const int32_t* src, int64_t len, void* dst, int32_t* esize);

/* Device-clock stopwatch (CUDA events on the context's stream).  hp_timer_stop returns the milliseconds since
 * hp_timer_start once the stream has drained.  Measurement aid (bench.py times its steps with it). */
int hp_timer_start(hp_ctx* ctx);
int hp_timer_stop(hp_ctx* ctx, float* ms);

/* ---- HiCCUPS scoring: callers.py:98-287 ------------------------------------------------------ */
typedef struct hp_hiccups_params {
    int32_t npw;                 /* number of (pw, ww) pairs                                   */
    int32_t pw[HP_MAX_PW];
    int32_t ww[HP_MAX_PW];
    int32_t maxww;
    int32_t min_local_reads;     /* callers.py:206                                             */
    int64_t maxapart_bins;       /* maxapart // res  (callers.py:102)                          */
    double sig;                  /* callers.py:273,279                                         */
    int32_t dump;                /* !=0: keep per-pixel bS/bE/E planes for hp_dump_plane (tests) */
    int32_t flags;               /* HP_PF_*                                                    */
} hp_hiccups_params;
#define HP_PF_BHFDR 2            /* the BH-FDR caller, callers.py:364-553: one pair, donut background only, Poisson
                                    rate = the pixel's own E, no lambda-chunks.  hp_hiccups_fdr then returns the
                                    pixels with p <= sig (flags REJECT_K, q = 1): the chromosome-wide BH over them and
                                    summary.lf[0][0].n_valid tests is finished by the caller (<= 1e5 records)      */
#define HP_PF_EXACT_SUMS 4       /* evaluate every pixel's sums in the reference's fp64 order (k_score_spec / k_score).
                                    Default for single-pair runs with maxww <= 10: the re-associated fp32 kernel
                                    classifies the pixels and only those it cannot settle (interval around E touches a
                                    lambda-chunk edge), E.max() and the survivors are evaluated in that order -- the
                                    results are bit-identical either way (tests/test_gpu_fast.py)                 */
#define HP_PF_GENERIC_KERNEL 1   /* use the table-driven score kernel even when a compiled-in sweep
                                    program matches (tests exercise both kernels)               */

typedef struct hp_step_stat {    /* one executed sweep step, callers.py:203-232                 */
    int32_t p, w;
    int64_t resolved;            /* 'Valid Contact Number from This Loop'                       */
    double valid_ratio, left_ratio;
} hp_step_stat;

typedef struct hp_lf_stat {      /* one (p, background) pair, callers.py:244-264                */
    int64_t n_valid;             /* pixels with E > 0                                           */
    double e_max;                /* E.max() (0 if n_valid == 0)                                 */
    int32_t numbin;              /* lambdachunk(): number of chunks                             */
    int32_t reserved;
    int64_t n_reject;            /* pixels with q <= sig (before the gap filter)                */
} hp_lf_stat;

typedef struct hp_hiccups_summary {
    int64_t n_pixels;            /* 'Observed Contact Number' (callers.py:114)                  */
    int64_t band_pixels;         /* dense band pixels min(ww) <= d <= maxapart_bins (the metric) */
    int32_t frozen_w;
    int32_t n_steps;             /* executed steps                                              */
    hp_step_stat steps[HP_MAX_STEPS];
    hp_lf_stat lf[HP_MAX_PW][2]; /* [pair][0 = donut 'K', 1 = lower-left 'Y']                    */
    int64_t n_candidates;
    int64_t n_survivors;         /* records available to hp_get_survivors                       */
    float ms_levels, ms_score, ms_fdr, ms_total; /* device time (CUDA events on the ctx stream): level kernel, score
                                                    kernel, BH + survivor kernels, their sum                      */
    int32_t launches;            /* kernels launched by this call                               */
    int32_t spec_kernel;         /* 1: a score kernel specialised for this sweep program ran     */
    float ms_exact;              /* device time of the exact re-evaluation after the re-associated kernel */
    int32_t fast_kernel;         /* 1: the re-associated kernel + exact re-evaluation ran (HP_PF_EXACT_SUMS off) */
    int64_t n_exact;             /* records the exact re-evaluation kernel took (0 without the fast kernel) */
} hp_hiccups_summary;

/* Host-only inspection hook (needs neither a GPU nor a context): the sweep program derived from prm->pw / ww / maxww
 * (callers.py:15-23, 132-198).  nsteps steps in execution order with their (p, w); step s adds the cells
 * [op_end[s-1], op_end[s]) of the op list, (opa, opb) = (row, column) offset in fp64 addition order, opy = the cell also
 * feeds the lower-left sum, opr = its raw count feeds Reads.  step_p / step_w / op_end need HP_MAX_STEPS entries; the op
 * arrays may be NULL to query *nops first. */
int hp_program_dump(const hp_hiccups_params* prm, int32_t* nsteps, int32_t* step_p, int32_t* step_w, int32_t* op_end,
                    int64_t op_capacity, int8_t* opa, int8_t* opb, uint8_t* opy, uint8_t* opr, int64_t* nops);

/* sweep (levels + adaptive width) + expected values + lambda-chunk histograms            */
int hp_hiccups_score(hp_ctx* ctx, const hp_hiccups_params* prm, hp_hiccups_summary* out);
/* Poisson + BH per lambda-chunk, survivor selection.  numbin_override: NULL, or [npw*2] chunk
 * counts to use instead of ceil(log(Emax)/log(2)*3+1) evaluated with the C library.        */
int hp_hiccups_fdr(hp_ctx* ctx, const int32_t* numbin_override, hp_hiccups_summary* out);
/* the summary of the last hp_hiccups_score as it stands now (after hp_allreduce_hist: genome-wide e_max / numbin / n_valid) */
int hp_get_summary(hp_ctx* ctx, hp_hiccups_summary* out);
/* both of the above */
int hp_hiccups(hp_ctx* ctx, const hp_hiccups_params* prm, hp_hiccups_summary* out);

typedef struct hp_survivor {     /* a pixel with q <= sig for K or Y of one pair                */
    int32_t r, c;                /* bin coordinates (x, y), c > r                               */
    int32_t pair;                /* index into pw/ww                                            */
    uint32_t flags;              /* HP_SF_*                                                     */
    double obs;                  /* raw count O                                                 */
    double ice;                  /* balanced value cM[r, c]                                     */
    double e[2];                 /* expected E for K, Y (0 if not valid)                        */
    double p[2];                 /* Poisson p                                                   */
    double q[2];                 /* BH q                                                        */
} hp_survivor;
#define HP_SF_VALID_K 1u         /* E_K > 0                                                     */
#define HP_SF_VALID_Y 2u
#define HP_SF_REJECT_K 4u        /* q_K <= sig                                                  */
#define HP_SF_REJECT_Y 8u
#define HP_SF_CEMY_NONZERO 16u   /* reference's cEM[ci,cj] != 0 for the Y background (callers.py:330) */
int hp_get_survivors(hp_ctx* ctx, hp_survivor* buf, int64_t capacity, int64_t* count);

/* ---- genome-wide FDR (optional; NOT the reference's behaviour, which corrects per chromosome) ------
 * Between hp_hiccups_score and hp_hiccups_fdr the (pair, background, lambda-chunk, observed) histograms can
 * be exported, summed over chromosomes / GPUs by the caller (one all-reduce), and imported back, so that BH
 * runs on the merged counts.  Layout: [npw * 2][total_bins] int64, total_bins from hp_hist_bins().       */
int hp_hist_bins(hp_ctx* ctx, int64_t* total_bins);
int hp_hist_export(hp_ctx* ctx, int64_t* out, int64_t capacity);
int hp_hist_import(hp_ctx* ctx, const int64_t* in, int64_t count);

/* The same merge on the device and across GPUs: the one collective of the path (replaces nothing in the reference, whose
 * dispatcher scripts/pyHICCUPS:192-198 gathers finished peak tables only).  hp_comm_init binds a context to rank `rank`
 * of an `nranks`-GPU NCCL communicator (one context per GPU; id = the HP_COMM_ID_BYTES bytes hp_comm_unique_id produced
 * on rank 0, passed to the other ranks by the caller; nranks == 1 needs no id and no NCCL).  hp_allreduce_hist, between
 * hp_hiccups_score and hp_hiccups_fdr: sums the u32 histograms of the `nctx` scored contexts of this process (all on
 * comm's GPU, same (pw, ww) list) into u64 on the device, ncclAllReduce(sum) over the ranks, max for E.max(), sum for
 * the valid counts, and writes the merged tables back into every context, whose summary (e_max, numbin, n_valid)
 * becomes the genome-wide one; hp_hiccups_fdr then runs BH on the merged counts.  Every rank must call it (nctx may
 * be 0) with the same npw = number of (pw, ww) pairs of the run.  *ms (optional): device time from the first local sum
 * to the last write-back. */
#define HP_COMM_ID_BYTES 128
int hp_comm_unique_id(void* id);
int hp_comm_init(hp_ctx* ctx, int32_t nranks, int32_t rank, const void* id);
int hp_comm_destroy(hp_ctx* ctx);
int hp_allreduce_hist(hp_ctx* comm, hp_ctx* const* ctxs, int32_t nctx, int32_t npw, float* ms);
/* gives back the upload scratch of a context (device landing zone, pinned staging); the uploaded band and the results
 * of the last score stay valid.  For callers that keep many scored contexts alive (one per chromosome until the merged BH). */
int hp_ctx_trim(hp_ctx* ctx);

/* rows of cM whose stored band is all zero ('gaps', callers.py:238): out[r] = 1 if row r is a gap
 * (available after hp_band_upload) */
int hp_get_gaps(hp_ctx* ctx, uint8_t* out, int64_t n);

/* ---- inspection (parity tests) --------------------------------------------------------------- */
/* per-pixel first sweep step at which Reads >= min_local_reads; 0xFE never, 0xFF not a pixel.
 * out is [num][n] row = diagonal. */
int hp_dump_levels(hp_ctx* ctx, uint8_t* out, int64_t capacity);
/* planes kept when params.dump != 0: what = 0 bS, 1 bE, 2 E ; out is [num][n] doubles, NaN = unset */
int hp_dump_plane(hp_ctx* ctx, int32_t pair, int32_t background, int32_t what, double* out, int64_t capacity);
/* lambda-chunk tables of one (pair, background): hist/p/q for chunk i in [1, numbin], bin k in
 * [0, width_i).  Call with NULL arrays to query sizes: widths[i-1] filled for i <= *numbin.     */
int hp_get_chunk_table(hp_ctx* ctx, int32_t pair, int32_t background, int32_t* numbin, int32_t* widths,
                       int64_t* hist, double* p, double* q, int64_t capacity);
/* device Poisson tail exactly as used for the tables: out[i] = 1 - pdtr(k[i], mu[i])           */
int hp_poisson_sf(hp_ctx* ctx, const double* k, const double* mu, double* out, int64_t count);

/* ---- APA pileup: /root/reference/hicpeaks/apa.py --------------------------------------------------- */
/* The balanced matrix of one chromosome as its upper diagonals 0 .. num-1 (bal_diags[d] has n - d doubles, NaN
 * kept: apa_submatrix rejects windows that hold a NaN, apa.py:20-22); the matrix is symmetric (cooler fetch). */
typedef struct hp_apa_desc {
    int64_t n;
    int32_t num;
    int32_t reserved;
    const double* const* bal_diags;   /* [num] */
} hp_apa_desc;
int hp_apa_upload(hp_ctx* ctx, const hp_apa_desc* band);
/* apa_submatrix (apa.py:11-28) + the per-window means of apa_analysis (apa.py:33): for anchor k the
 * (2w+1)^2 window around (pos_i[k], pos_j[k]) is gathered, rejected (valid[k] = 0) when it leaves the matrix,
 * holds a NaN or has mean 0, else divided by its mean and kept on the device; mean_arr[k] = mean of the
 * normalised window, bit-identical to numpy's.  Every anchor needs |j - i| + 2w < num. */
int hp_apa_windows(hp_ctx* ctx, const int32_t* pos_i, const int32_t* pos_j, int64_t npos, int32_t w, uint8_t* valid,
                   double* mean_arr);
/* the same means for caller-supplied normalised windows (apa_analysis on a host array): wins is
 * npos * (2w+1)^2 doubles; the windows stay on the device for hp_apa_accumulate */
int hp_apa_load_windows(hp_ctx* ctx, const double* wins, int64_t npos, int32_t w, double* mean_arr);
/* the axis-0 sum of apa[mask].mean(axis=0) (apa.py:37): acc[(2w+1)^2] = (init ? 0 : acc) + windows sel[0..nsel)
 * added one by one in that order (numpy's order); the caller divides by the total count */
int hp_apa_accumulate(hp_ctx* ctx, const int64_t* sel, int64_t nsel, double* acc, int32_t init);
/* download normalised windows (apa_submatrix's return value), out is nsel * (2w+1)^2 doubles */
int hp_apa_get_windows(hp_ctx* ctx, const int64_t* sel, int64_t nsel, double* out);

#ifdef __cplusplus
}
#endif
#endif /* HICPEAKS_B200_H */
