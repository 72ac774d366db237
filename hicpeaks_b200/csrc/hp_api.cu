// Host side of the C ABI declared in include/hicpeaks_b200.h: context, band upload, the sweep
// program builder (callers.py:15-23,132-198), the frozen_w replay (callers.py:203-232) and the
// launch sequence  K1 levels -> replay -> bE table -> K2 score -> K3 BH -> survivor filter.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "hp_kernels.cuh"
#include "hp_score_spec.cuh"
#include "hp_score_fast.cuh"
#include "hp_exact.cuh"
#include "hp_apa.cuh"
#include "hp_prep.cuh"
#include "hp_hostpack.h"

using namespace hp;

#ifndef HP_SYNC_DEFAULT
#define HP_SYNC_DEFAULT 2
#endif

static thread_local std::string g_err;

struct hp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    int sm_count = 148;
    cudaEvent_t ev_sync = nullptr;    // blocking-sync event (stream_sync)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;      // hp_timer_start / hp_timer_stop
    std::string err;
    // sweep program + chunk tables: pinned host mirror and the device copy the kernels read
    Tables* h_tab = nullptr;
    Tables* d_tab = nullptr;
    Chunks chunks{};
    std::vector<double> h_ptab;
    double* d_ptab = nullptr;
    // band
    int64_t n = 0;
    int num = 0, bal_first = 0, pitch = 0;
    size_t plane = 0;                 // num * pitch
    size_t cap_plane = 0, cap_n = 0;
    int* d_raw = nullptr;
    double* d_bal = nullptr;
    unsigned char* d_lvl = nullptr;
    double *d_ir = nullptr, *d_b1 = nullptr, *d_b2 = nullptr;
    unsigned int* d_rownz = nullptr;
    double* d_tmp = nullptr;          // plain-layout landing zone of the uploads (re-laid out on the device)
    void* h_stage = nullptr;          // pinned staging for uploads
    size_t cap_stage = 0;
    unsigned char* h_res = nullptr;   // pinned, 8 KB: landing zone of the small per-call readbacks (level histogram, counters).
                                      // A D2H copy into pageable memory blocks inside the copy call; into pinned memory it
                                      // is queued and the thread waits in stream_sync, the way HP_SYNC says
    void* h_out = nullptr;            // pinned staging for result downloads (a D2H copy into pageable memory runs at ~3 GB/s)
    size_t cap_out = 0;
    bool have_band = false;
    // run state
    hp_hiccups_params prm{};
    Prog prog{};
    std::vector<signed char> opa, opb;
    std::vector<unsigned char> opy, opr;
    std::vector<unsigned short> ropi;
    bool scored = false, fdr_done = false;
    bool kcand_valid = false, kcand_bhfdr = false;     // candidate thresholds cached per (sig, mode)
    double kcand_sig = 0.0;
    hp_hiccups_summary sum{};
    int dlo = 0, dhi = -1;
    unsigned long long* d_lhist = nullptr;      // [HP_MAX_STEPS + 2]
    double* d_betab = nullptr; size_t cap_betab = 0;
    unsigned int* d_hist = nullptr; size_t cap_hist = 0;
    double* d_qtab = nullptr; size_t cap_qtab = 0;
    unsigned long long* d_small = nullptr;      // [0..15] emax bits, [16..31] nvalid, [32..47] nreject
    unsigned int* d_cnt = nullptr;              // [0..3] cand counters, [4..7] survivor counters, [16..31] fast kernel: running E.max() bounds
    int* d_numbin = nullptr;                    // [16]
    int* d_kq = nullptr;                        // [16][kMaxChunk + 2] per (pair, background, chunk): smallest count with q <= sig (k_bh)
    Cand* d_cand = nullptr; size_t cap_cand = 0;
    // re-associated score kernel (hp_score_fast.cuh): classified candidates without E, records for k_exact, fp32 factors
    FCand* d_fcand = nullptr; size_t cap_fcand = 0;
    XRec* d_xrec = nullptr; size_t cap_xrec = 0;
    float* d_ffac = nullptr; size_t cap_ffac = 0;
    float* d_ffs = nullptr; size_t cap_ffs = 0;        // per-strip blocks of the interior factors (bulk-copied by the fast kernel)
    float* d_xf = nullptr; size_t cap_xf = 0;          // fp32 row-major copy of the balanced band (k_f32plane), row pitch xf_ndp
    float *d_b1f = nullptr, *d_b2s = nullptr; size_t cap_b1f = 0, cap_b2s = 0;
    int xf_ndp = 0;
    unsigned int nfcand = 0;
    bool fast_used = false;
    bool edges_regular = false;           // every chunk edge is 2^e times rv[2] or rv[3]: the fast kernel's mantissa compares apply
    bool domain_ok = false;               // every balanced value is inside the domain of the fast kernel's error bound
    // genome-wide FDR (hp_comm_init / hp_allreduce_hist): NCCL communicator + the u64 accumulator of the histograms
    ncclComm_t comm = nullptr;
    int comm_nranks = 0, comm_rank = 0;
    unsigned long long* d_acc = nullptr; size_t cap_acc = 0;
    int4* d_fscratch = nullptr;           // [sm_count][kFScratch] E.max() contenders of the fast kernel's CTAs
    hp_survivor* d_surv = nullptr; size_t cap_surv = 0;
    double* d_dump = nullptr; size_t cap_dump = 0;
    int numbin[HP_MAX_PW * 2] = {};
    unsigned int ncand = 0, nsurv = 0;
    bool spec_used = false;
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    // K0 scratch (hp_prep.cuh)
    double* d_w = nullptr; size_t cap_w = 0;
    PackedDiag* d_pk = nullptr; size_t cap_pk = 0;       // per-diagonal format table of the narrowed count upload
    int64_t h2d_bytes = 0;                               // bytes the last band upload moved over PCIe
    unsigned char* d_prep = nullptr; size_t cap_prep = 0;
    // APA (hp_apa.cuh)
    double* d_apa_bal = nullptr; size_t cap_apa_bal = 0;
    int64_t apa_n = 0; int apa_num = 0;
    ApaPlan* d_apa_plan = nullptr;
    int* d_apa_pos = nullptr; size_t cap_apa_pos = 0;
    double* d_apa_wins = nullptr; size_t cap_apa_wins = 0;
    unsigned char* d_apa_valid = nullptr; size_t cap_apa_valid = 0;
    double* d_apa_mean = nullptr; size_t cap_apa_mean = 0;
    long long* d_apa_sel = nullptr; size_t cap_apa_sel = 0;
    double* d_apa_avg = nullptr; size_t cap_apa_avg = 0;
    int64_t apa_npos = 0; int apa_w = 0;
};

// How a host thread waits for its context's stream (HP_SYNC=spin|yield|block):
//   spin   cudaStreamSynchronize, the runtime's default policy: lowest latency while every waiting thread has a core
//   yield  poll cudaStreamQuery and give the core away between polls: as quick as spinning on an idle box, and the
//          waiting threads (one per chromosome in flight, times the ranks on the box) do not starve the ones that
//          have launches to issue when they outnumber the cores
//   block  a blocking-sync event: the thread sleeps; highest wake-up latency
//   hybrid yield-poll for HP_SYNC_SPIN_US (40) microseconds, then sleep on the blocking event: short waits stay quick,
//          long ones (a score kernel queued behind seven others) leave the core to the threads packing uploads
static int sync_mode() {
    static const int mode = []() {
        const char* e = getenv("HP_SYNC");
        if (e && !strcmp(e, "spin")) return 0;
        if (e && !strcmp(e, "block")) return 1;
        if (e && !strcmp(e, "yield")) return 2;
        if (e && !strcmp(e, "hybrid")) return 3;
        return HP_SYNC_DEFAULT;
    }();
    return mode;
}
static int hybrid_spin_us() {
    static const int us = []() { const char* e = getenv("HP_SYNC_SPIN_US"); return e ? std::max(0, atoi(e)) : 40; }();
    return us;
}
static cudaError_t stream_sync(hp_ctx* ctx) {
    const int m = sync_mode();
    if (m == 2) {
        for (;;) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaErrorNotReady) return q;
            sched_yield();
        }
    }
    if (m == 3 && ctx->ev_sync) {                      // hybrid: poll for a short while, then sleep on a blocking event
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaErrorNotReady) return q;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(hybrid_spin_us())) break;
            sched_yield();
        }
        cudaError_t e = cudaEventRecord(ctx->ev_sync, ctx->stream);
        return e == cudaSuccess ? cudaEventSynchronize(ctx->ev_sync) : e;
    }
    if (m == 0 || !ctx->ev_sync) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->ev_sync, ctx->stream);
    return e == cudaSuccess ? cudaEventSynchronize(ctx->ev_sync) : e;
}

static int fail(hp_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    g_err = msg;
    return code;
}
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, HP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when a launch needs more than was granted before
// (per function and device; the attribute sticks)
template <typename K>
static cudaError_t want_smem(K kernel, int device, size_t smem, std::atomic<size_t>* granted) {
    std::atomic<size_t>& g = granted[device & 63];
    if (smem <= g.load(std::memory_order_acquire)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) g.store(smem, std::memory_order_release);
    return e;
}

static cudaError_t ensure_host(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return cudaSuccess;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    bytes = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

template <typename T>
static cudaError_t ensure(T** p, size_t* cap, size_t want) {
    if (*cap >= want && *p) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    want += want / 4;                                  // slack: a slightly larger request next call does not reallocate
    cudaError_t e = cudaMalloc((void**)p, want * sizeof(T));
    if (e == cudaSuccess) *cap = want;
    return e;
}

extern "C" int hp_abi_version(void) { return HP_ABI_VERSION; }

extern "C" int hp_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(nullptr, HP_ERR_NO_DEVICE, cudaGetErrorString(e)); }
    *count = n;
    return HP_OK;
}

extern "C" const char* hp_last_error(const hp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

static int build_chunks(hp_ctx* ctx, int max_chunks, const double* edges) {
    Chunks& C = ctx->chunks;
    memset(&C, 0, sizeof(C));
    C.maxchunk = max_chunks;
    C.rv[0] = 0.0;
    int off = 0;
    for (int i = 1; i <= max_chunks; ++i) {
        C.rv[i] = edges ? edges[i - 1] : (i == 1 ? 1.0 : pow(2.0, (i - 1) / 3.0));
        if (!(C.rv[i] > C.rv[i - 1])) return fail(ctx, HP_ERR_INVALID, "chunk edges must increase");
        // the kernels locate a chunk from the binary exponent: every third edge must be 2^((i-1)/3) exactly
        if (i % 3 == 1 && C.rv[i] != ldexp(1.0, (i - 1) / 3))
            return fail(ctx, HP_ERR_INVALID, "chunk edge " + std::to_string(i) + " is not the exact power of two 2^((i-1)/3)");
        C.hoff[i] = off;
        C.hw[i] = (int)ceil(C.rv[i] + 12.0 * sqrt(C.rv[i]) + 40.0);
        off += C.hw[i];
    }
    C.hoff[max_chunks + 1] = off;
    C.total_bins = off;
    // the re-associated kernel finds a chunk from the exponent and two mantissa compares: edges 3e+2, 3e+3 must be 2^e times
    // the in-octave edges rv[2], rv[3] (to 1e-12; the kernel's thresholds carry a 1e-9 margin)
    ctx->edges_regular = max_chunks >= 3;
    for (int i = 2; i <= max_chunks && ctx->edges_regular; ++i) {
        if (i % 3 == 1) continue;
        const double ref = ldexp(C.rv[i % 3 == 2 ? 2 : 3], (i - 2) / 3);
        if (!(fabs(C.rv[i] / ref - 1.0) < 1e-12)) ctx->edges_regular = false;
    }
    return HP_OK;
}

extern "C" int hp_ctx_create(int device, int max_chunks, const double* edges, hp_ctx** out) {
    if (!out) return fail(nullptr, HP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (max_chunks == 0) max_chunks = 52;
    if (max_chunks < 1 || max_chunks > kMaxChunk) return fail(nullptr, HP_ERR_INVALID, "max_chunks out of range [1,64]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, HP_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, HP_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, HP_ERR_NO_DEVICE, "device is not sm_100 (kernels are built for sm_100a only)");
    hp_ctx* ctx = new hp_ctx();
    ctx->device = device;
    auto bail = [&](int code) { hp_ctx_destroy(ctx); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(fail(nullptr, HP_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(nullptr, HP_ERR_CUDA, "stream create failed"));
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming);
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
        return bail(fail(nullptr, HP_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver"));
    ctx->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    int rc = build_chunks(ctx, max_chunks, edges);
    if (rc) { g_err = ctx->err; return bail(rc); }
    const size_t tb = ctx->chunks.total_bins;
    bool ok = cudaMalloc(&ctx->d_ptab, tb * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&ctx->d_tab, sizeof(Tables)) == cudaSuccess &&
              cudaHostAlloc(&ctx->h_tab, sizeof(Tables), cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc((void**)&ctx->h_res, 8192, cudaHostAllocDefault) == cudaSuccess &&
              cudaMalloc(&ctx->d_lhist, (HP_MAX_STEPS + 2) * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&ctx->d_small, 48 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&ctx->d_cnt, 48 * sizeof(unsigned int)) == cudaSuccess &&
              cudaMalloc(&ctx->d_numbin, 16 * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&ctx->d_kq, 16 * (kMaxChunk + 2) * sizeof(int)) == cudaSuccess;
    if (!ok) return bail(fail(nullptr, HP_ERR_CUDA, "device allocation failed"));
    // Poisson table (universal): p[i][k] = 1 - pdtr(k, rv_i)
    for (int i = 1; i <= max_chunks; ++i) ctx->chunks.kcand[i] = 0;
    memset(ctx->h_tab, 0, sizeof(Tables));
    ctx->h_tab->chunks = ctx->chunks;
    if (cudaMemcpyAsync(ctx->d_tab, ctx->h_tab, sizeof(Tables), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return bail(fail(nullptr, HP_ERR_CUDA, "table upload failed"));
    k_ptab<<<(unsigned)((tb + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_tab, ctx->d_ptab);
    ctx->h_ptab.resize(tb);
    cudaMemcpyAsync(ctx->h_ptab.data(), ctx->d_ptab, tb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = stream_sync(ctx);
    if (e != cudaSuccess) return bail(fail(nullptr, HP_ERR_CUDA, std::string("Poisson table kernel: ") + cudaGetErrorString(e)));
    for (int i = 1; i <= max_chunks; ++i)
        if (ctx->h_ptab[ctx->chunks.hoff[i] + ctx->chunks.hw[i] - 1] != 0.0)
            return bail(fail(nullptr, HP_ERR_INVALID, "internal: Poisson table too narrow for chunk " + std::to_string(i)));
    *out = ctx;
    return HP_OK;
}

static void comm_release(hp_ctx* ctx);
extern "C" void hp_ctx_destroy(hp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) stream_sync(ctx);
    comm_release(ctx);
    void* ptrs[] = {ctx->d_ptab, ctx->d_raw, ctx->d_bal, ctx->d_lvl, ctx->d_ir, ctx->d_b1, ctx->d_b2, ctx->d_rownz,
                    ctx->d_lhist, ctx->d_betab, ctx->d_hist, ctx->d_qtab, ctx->d_small, ctx->d_cnt, ctx->d_numbin, ctx->d_kq,
                    ctx->d_cand, ctx->d_fcand, ctx->d_xrec, ctx->d_ffac, ctx->d_ffs, ctx->d_xf, ctx->d_b1f, ctx->d_b2s, ctx->d_fscratch, ctx->d_acc, ctx->d_surv, ctx->d_dump, ctx->d_tmp, ctx->d_tab, ctx->d_w, ctx->d_pk, ctx->d_prep, ctx->d_apa_bal, ctx->d_apa_plan, ctx->d_apa_pos,
                    ctx->d_apa_wins, ctx->d_apa_valid, ctx->d_apa_mean, ctx->d_apa_sel, ctx->d_apa_avg};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_tab) cudaFreeHost(ctx->h_tab);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---------------------------------------------------------------------------------------------
// Host threads that pack diagonals into the pinned staging buffers: ONE pool per process, shared by every
// context, so that eight chromosomes uploading at once (one caller thread per context) split the cores between
// them instead of starting eight private thread teams.  HP_PACK_THREADS overrides the pool size.
class PackPool {
  public:
    static PackPool& get() {
        static PackPool* p = new PackPool();          // never destroyed: workers may outlive static destructors
        return *p;
    }
    unsigned size() const { return (unsigned)workers_.size(); }
    // Runs fn(0) .. fn(count - 1) on the pool; calls ready(k) on the CALLING thread, in order k = 0, 1, ..., as soon
    // as fn(k) has finished (the caller issues chunk k's copy while later chunks are still being packed).
    template <typename F, typename R>
    void run_ordered(int count, F&& fn, R&& ready) {
        if (count <= 0) return;
        if (workers_.empty() || count == 1 || forked_) {
            for (int k = 0; k < count; ++k) { fn(k); ready(k); }
            return;
        }
        Job job;
        job.count = count;
        job.done.assign(count, 0);
        job.fn = [&fn](int k) { fn(k); };
        {
            std::lock_guard<std::mutex> lk(m_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        for (int k = 0; k < count; ++k) {
            for (;;) {
                int mine = -1;
                {
                    std::unique_lock<std::mutex> lk(m_);
                    if (job.done[k]) break;
                    if (job.next < job.count) mine = job.next++;         // the caller packs too instead of idling
                    else { job.cv.wait(lk, [&] { return job.done[k] != 0; }); break; }
                }
                fn(mine);
                std::lock_guard<std::mutex> lk(m_);
                job.done[mine] = 1;
                ++job.finished;
            }
            ready(k);
        }
        std::unique_lock<std::mutex> lk(m_);                              // workers may still be leaving the job
        job.cv.wait(lk, [&] { return job.finished == job.count && job.active == 0; });
        jobs_.erase(std::find(jobs_.begin(), jobs_.end(), &job));
    }

  private:
    struct Job {
        int count = 0, next = 0, finished = 0, active = 0;
        std::vector<char> done;
        std::function<void(int)> fn;
        std::condition_variable cv;
    };
    PackPool() {
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) hw = std::max(2u, hw / (unsigned)std::max(1, atoi(e)));   // ranks share the box
        // callers pack too (one per chromosome in flight): half the cores for the pool keeps callers + workers at about
        // one thread per core (16-core box, 8 chromosomes in flight: 4.9 ms per step against 5.4 with 16 pool threads)
        unsigned n = std::min<unsigned>(16u, std::max(2u, hw / 2));
        if (const char* e = getenv("HP_PACK_THREADS")) n = (unsigned)std::max(1, atoi(e));
        for (unsigned t = 0; t + 1 < n; ++t) workers_.emplace_back([this] { work(); }), workers_.back().detach();
        pthread_atfork(nullptr, nullptr, [] { forked_ = true; });      // a forked child has no workers: pack inline
    }
    void work() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            Job* job = nullptr;
            for (Job* j : jobs_)
                if (j->next < j->count) { job = j; break; }
            if (!job) { cv_.wait(lk); continue; }
            const int k = job->next++;
            ++job->active;
            lk.unlock();
            job->fn(k);
            lk.lock();
            job->done[k] = 1;
            ++job->finished;
            --job->active;
            job->cv.notify_all();
        }
    }
    static inline bool forked_ = false;
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<Job*> jobs_;
    std::vector<std::thread> workers_;
};

// Row pitch of the band planes: a multiple of 64, so that a sub-plane of the u8 level plane (pitch / 4 bytes) starts on a
// 16-byte boundary -- TMA strides and box origins must (hp_score_fast.cuh loads level tiles as 3-D boxes)
static inline int band_pitch(int64_t n) { return (int)((n + 63) / 64 * 64); }

// what the re-associated score kernel reads besides the planes: the fp32 row-major copy of the balanced band and the
// fp32 bias vectors (hp_score_fast.cuh); queued on the context's stream right after the planes have been written
static int fast_companions(hp_ctx* ctx, int64_t n, int num, int bf, int pitch) {
    const int ndp = (num - bf + 3) & ~3;
    const int n1 = (int)((n + kFTR - 1) / kFTR * kFTR) + kFTR, n2 = (int)n + num + 4 * kFTD;
    CK(ensure(&ctx->d_xf, &ctx->cap_xf, (size_t)n * ndp));
    CK(ensure(&ctx->d_b1f, &ctx->cap_b1f, (size_t)n1));
    CK(ensure(&ctx->d_b2s, &ctx->cap_b2s, (size_t)n2));
    k_f32plane<<<dim3((unsigned)((n + 63) / 64), (unsigned)((ndp + 31) / 32)), 256, 0, ctx->stream>>>(ctx->d_bal, ctx->d_xf, (int)n, num, bf, pitch, ndp);
    CK(cudaGetLastError());
    k_fast_bias<<<(unsigned)((std::max(n1, n2) + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_b1, ctx->d_b2, ctx->d_b1f, ctx->d_b2s, (int)n, bf, n1, n2);
    CK(cudaGetLastError());
    ctx->xf_ndp = ndp;
    return HP_OK;
}

static int band_alloc(hp_ctx* ctx, int64_t n_, int num_, int bal_first_) {
    struct { int64_t n; int num; int bal_first; } bb{n_, num_, bal_first_};
    auto* b = &bb;
    if (!ctx) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (b->n <= 0 || b->n > (1ll << 30) || b->num <= 0 || b->num > b->n || b->bal_first < 0 || b->bal_first >= b->num)
        return fail(ctx, HP_ERR_INVALID, "bad band geometry (need 0 < num <= n, 0 <= bal_first < num)");
    CK(cudaSetDevice(ctx->device));
    ctx->have_band = false; ctx->scored = false; ctx->fdr_done = false;
    const int64_t n = b->n;
    const int num = b->num;
    const int pitch = band_pitch(n);
    const size_t plane = (size_t)num * pitch;
    if (plane > ctx->cap_plane) {
        for (void* p : {(void*)ctx->d_raw, (void*)ctx->d_bal, (void*)ctx->d_lvl, (void*)ctx->d_tmp}) if (p) cudaFree(p);
        ctx->d_raw = nullptr; ctx->d_bal = nullptr; ctx->d_lvl = nullptr; ctx->d_tmp = nullptr; ctx->cap_plane = 0;
        CK(cudaMalloc(&ctx->d_raw, plane * sizeof(int)));
        CK(cudaMalloc(&ctx->d_bal, plane * sizeof(double)));
        CK(cudaMalloc(&ctx->d_tmp, plane * 12));
        CK(cudaMalloc(&ctx->d_lvl, plane));
        ctx->cap_plane = plane;
    }
    if ((size_t)n > ctx->cap_n) {
        for (void* p : {(void*)ctx->d_ir, (void*)ctx->d_b1, (void*)ctx->d_b2, (void*)ctx->d_rownz}) if (p) cudaFree(p);
        ctx->d_ir = ctx->d_b1 = ctx->d_b2 = nullptr; ctx->d_rownz = nullptr; ctx->cap_n = 0;
        CK(cudaMalloc(&ctx->d_ir, n * sizeof(double)));
        CK(cudaMalloc(&ctx->d_b1, n * sizeof(double)));
        CK(cudaMalloc(&ctx->d_b2, n * sizeof(double)));
        CK(cudaMalloc(&ctx->d_rownz, n * sizeof(unsigned int)));
        ctx->cap_n = n;
    }
    const size_t stage_bytes = plane * 12 + (size_t)num * 8 + (size_t)num * sizeof(PackedDiag);
    if (stage_bytes > ctx->cap_stage) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->cap_stage = 0;
        CK(cudaHostAlloc(&ctx->h_stage, stage_bytes, cudaHostAllocDefault));
        ctx->cap_stage = stage_bytes;
    }
    return HP_OK;
}

extern "C" int hp_band_upload(hp_ctx* ctx, const hp_band_desc* b) {
    if (!ctx || !b) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!b->raw_diags || !b->bal_diags || !b->ir || !b->b1 || !b->b2) return fail(ctx, HP_ERR_INVALID, "NULL band array");
    {
        const int rc0 = band_alloc(ctx, b->n, b->num, b->bal_first);
        if (rc0) return rc0;
    }
    const int64_t n = b->n;
    const int num = b->num;
    const int pitch = band_pitch(n);
    const size_t plane = (size_t)num * pitch;
    double* hbal = (double*)ctx->h_stage;
    int* hraw = (int*)((char*)ctx->h_stage + plane * 8);
    double* hir = (double*)((char*)ctx->h_stage + plane * 12);
    // pack the diagonals into the pitched planes (zero tails), a few host threads
    const int bf = b->bal_first;
    auto pack = [&](int d_begin, int d_end) {
        for (int d = d_begin; d < d_end; ++d) {
            const size_t len = (size_t)(n - d);
            int* rr = hraw + (size_t)d * pitch;
            stream_copy(rr, b->raw_diags[d], len * sizeof(int));
            memset(rr + len, 0, (pitch - len) * sizeof(int));
            double* br = hbal + (size_t)d * pitch;
            if (d >= bf) {
                stream_copy(br, b->bal_diags[d - bf], len * sizeof(double));
                memset(br + len, 0, (pitch - len) * sizeof(double));
            } else {
                memset(br, 0, (size_t)pitch * sizeof(double));
            }
        }
    };
    // chunks of diagonals: worker threads pack chunk k + 1 .. while the copy engine uploads chunk k
    for (int d = 0; d < num; ++d) hir[d] = d >= bf ? b->ir[d - bf] : 0.0;
    double* tbal = ctx->d_tmp;
    int* traw = (int*)((char*)ctx->d_tmp + plane * 8);
    const int per = std::max(1, (int)((size_t)(4u << 20) / ((size_t)pitch * 12)));     // ~4 MB per chunk
    const int nchunk = (num + per - 1) / per;
    cudaError_t cerr = cudaSuccess;
    auto send = [&](int k) {
        const int d0 = k * per, d1 = std::min(num, d0 + per);
        const size_t off = (size_t)d0 * pitch, cnt = (size_t)(d1 - d0) * pitch;
        cudaError_t e = cudaMemcpyAsync(tbal + off, hbal + off, cnt * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(traw + off, hraw + off, cnt * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess && cerr == cudaSuccess) cerr = e;
    };
    PackPool::get().run_ordered(nchunk, [&](int k) { pack(k * per, std::min(num, (k + 1) * per)); }, send);
    if (cerr != cudaSuccess) return fail(ctx, HP_ERR_CUDA, std::string("band upload: ") + cudaGetErrorString(cerr));
    CK(cudaMemsetAsync(ctx->d_rownz, 0, (size_t)n * sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_cnt + 12, 0, sizeof(unsigned int), ctx->stream));
    k_relayout<double><<<dim3((pitch + 255) / 256, num), 256, 0, ctx->stream>>>(tbal, ctx->d_bal, ctx->d_rownz, pitch, num, ctx->d_cnt + 12);
    CK(cudaGetLastError());
    k_relayout<int><<<dim3((pitch + 255) / 256, num), 256, 0, ctx->stream>>>(traw, ctx->d_raw, nullptr, pitch, num, nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->d_ir, hir, (size_t)num * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_b1, b->b1, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_b2, b->b2, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    { const int rcf = fast_companions(ctx, n, num, bf, pitch); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(ctx->h_res + 6144, ctx->d_cnt + 12, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(stream_sync(ctx));
    ctx->domain_ok = *(const unsigned int*)(ctx->h_res + 6144) == 0u;
    ctx->n = n; ctx->num = num; ctx->bal_first = bf; ctx->pitch = pitch; ctx->plane = plane;
    ctx->h2d_bytes = (int64_t)(plane * 12 + (size_t)num * 8 + (size_t)n * 16);
    ctx->have_band = true;
    return HP_OK;
}

// Device-clock stopwatch on the context's stream: a CUDA event now, a second one at stop, elapsed time between them.
// bench.py brackets a whole step with it (every context of the step has been synchronised before stop is called), so
// the step is timed on the device, not by the host clock.
extern "C" int hp_timer_start(hp_ctx* ctx) {
    if (!ctx) return fail(ctx, HP_ERR_INVALID, "NULL ctx");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->ev_t0) { CK(cudaEventCreate(&ctx->ev_t0)); CK(cudaEventCreate(&ctx->ev_t1)); }
    CK(cudaEventRecord(ctx->ev_t0, ctx->stream));
    return HP_OK;
}
extern "C" int hp_timer_stop(hp_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->ev_t0) return fail(ctx, HP_ERR_STATE, "hp_timer_start must come first");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev_t1));
    CK(cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
    return HP_OK;
}

// host-only inspection hook (no GPU, no context): the narrowing hp_band_upload_counts applies to one count diagonal
extern "C" int hp_narrow_diagonal(const int32_t* src, int64_t len, void* dst, int32_t* esize) {
    if ((!src && len > 0) || !dst || !esize || len < 0) return fail(nullptr, HP_ERR_INVALID, "NULL argument");
    *esize = narrow_diagonal(src, (size_t)len, dst);
    return HP_OK;
}

extern "C" int hp_upload_bytes(hp_ctx* ctx, int64_t* bytes) {
    if (!ctx || !bytes) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->have_band) return fail(ctx, HP_ERR_STATE, "no band uploaded");
    *bytes = ctx->h2d_bytes;
    return HP_OK;
}

// ---------------------------------------------------------------------------------------------
// worker-level entry: raw counts + balancing weights; the balanced band, IR and the biases are built on the
// device exactly as the reference's worker builds them on the host (scripts/pyHICCUPS:143-166)
extern "C" int hp_band_upload_counts(hp_ctx* ctx, const hp_counts_desc* b) {
    if (!ctx || !b) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!b->raw_diags || !b->weights) return fail(ctx, HP_ERR_INVALID, "NULL band array");
    {
        const int rc0 = band_alloc(ctx, b->n, b->num, b->bal_first);
        if (rc0) return rc0;
    }
    const int64_t n = b->n;
    const int num = b->num, bf = b->bal_first;
    const int pitch = band_pitch(n);
    const size_t plane = (size_t)num * pitch;
    cudaStream_t st = ctx->stream;
    int* hraw = (int*)((char*)ctx->h_stage + plane * 8);
    double* comp = ctx->d_tmp;                        // [num - bf][pitch] compaction scratch (the unused balanced landing zone)
    CK(ensure(&ctx->d_w, &ctx->cap_w, (size_t)n + 64));
    int depth = 0;                                     // levels of numpy's pairwise recursion below the root (hp_prep.cuh)
    while ((pitch >> depth) + 16 > 128) ++depth;   // a node of level l holds <= m / 2^l + 15 values
    const int nslot = 2 << depth;
    const int nb = num - bf;
    const size_t prep_bytes = (size_t)nb * nslot * (8 + 8 + 4);
    CK(ensure(&ctx->d_prep, &ctx->cap_prep, prep_bytes));
    int2* leaf = (int2*)ctx->d_prep;
    double* tval = (double*)(leaf + (size_t)nb * nslot);
    int* tlist = (int*)(tval + (size_t)nb * nslot);
    CK(cudaMemcpyAsync(ctx->d_w, b->weights, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    // raw counts: worker threads narrow each diagonal to u8 / u16 / i32 (hp_hostpack.cpp) straight into the pinned
    // staging buffer, chunk by chunk, while the copy engine uploads the chunks already done; k_prep_band reads
    // the narrowed form directly.  A chunk owns the byte range its diagonals
    // would take as int32, so chunks are packed independently and only the bytes used are sent.
    unsigned char* hpk = (unsigned char*)hraw;                       // pinned, plane * 4 bytes
    unsigned char* dpk = (unsigned char*)ctx->d_tmp + plane * 8;     // device: the int32 part of the landing zone (comp owns the first plane * 8 bytes)
    PackedDiag* htab = (PackedDiag*)((char*)ctx->h_stage + plane * 12 + (size_t)num * 8);
    CK(ensure(&ctx->d_pk, &ctx->cap_pk, (size_t)num));
    const int per = std::max(1, (int)((size_t)(5u << 19) / ((size_t)pitch * 4)));      // ~2.5 MB of int32 per chunk
    const int nchunk = (num + per - 1) / per;
    std::vector<size_t> used(nchunk);
    auto pack = [&](int k) {
        const int d0 = k * per, d1 = std::min(num, d0 + per);
        const size_t base = (size_t)d0 * pitch * 4;
        size_t off = base;
        for (int d = d0; d < d1; ++d) {
            const size_t len = (size_t)(n - d);
            const int es = narrow_diagonal((const int32_t*)b->raw_diags[d], len, hpk + off);
            htab[d].off = off; htab[d].esize = es; htab[d].len = (unsigned)len;
            off += (len * es + 15) & ~(size_t)15;
        }
        used[k] = off - base;
    };
    cudaError_t cerr = cudaSuccess;
    size_t sent = 0;
    auto send = [&](int k) {
        const size_t base = (size_t)k * per * pitch * 4;
        cudaError_t e = cudaMemcpyAsync(dpk + base, hpk + base, used[k], cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess && cerr == cudaSuccess) cerr = e;
        sent += used[k];
    };
    static const bool trace = getenv("HP_TRACE") != nullptr;          // phase times of the upload on stderr (development aid)
    const auto t_begin = std::chrono::steady_clock::now();
    PackPool::get().run_ordered(nchunk, pack, send);
    const auto t_packed = std::chrono::steady_clock::now();
    if (cerr != cudaSuccess) return fail(ctx, HP_ERR_CUDA, std::string("band upload: ") + cudaGetErrorString(cerr));
    CK(cudaMemcpyAsync(ctx->d_pk, htab, (size_t)num * sizeof(PackedDiag), cudaMemcpyHostToDevice, st));
    ctx->h2d_bytes = (int64_t)(sent + (size_t)num * sizeof(PackedDiag) + (size_t)n * 8);
    CK(cudaMemsetAsync(ctx->d_rownz, 0, (size_t)n * sizeof(unsigned int), st));
    CK(cudaMemsetAsync(ctx->d_ir, 0, (size_t)num * 8, st));
    CK(cudaMemsetAsync(ctx->d_cnt + 12, 0, sizeof(unsigned int), st));
    if (bf > 0) {
        const size_t cnt = (size_t)bf * pitch;
        k_zero_planes<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(ctx->d_bal, cnt);
        k_unpack_quad<<<dim3((pitch / 4 + 255) / 256, bf), 256, 0, st>>>(dpk, ctx->d_pk, ctx->d_raw, pitch);
        CK(cudaGetLastError());
    }
    const int prep_smem = nslot * 20 <= 40 * 1024 ? nslot * 20 : 0;
    k_prep_band<<<nb, kPrepThreads, prep_smem, st>>>(dpk, ctx->d_pk, ctx->d_w, (int)n, num, pitch, bf, ctx->d_raw, ctx->d_bal, ctx->d_rownz,
                                                   ctx->d_ir, comp, leaf, tval, tlist, depth, nslot, prep_smem ? 1 : 0, ctx->d_cnt + 12);
    CK(cudaGetLastError());
    k_prep_bias<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->d_w, ctx->d_b1, ctx->d_b2, (int)n);
    CK(cudaGetLastError());
    { const int rcf = fast_companions(ctx, n, num, bf, pitch); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(ctx->h_res + 6144, ctx->d_cnt + 12, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    const auto t_launched = std::chrono::steady_clock::now();
    CK(stream_sync(ctx));
    ctx->domain_ok = *(const unsigned int*)(ctx->h_res + 6144) == 0u;
    if (trace) {
        const auto t_end = std::chrono::steady_clock::now();
        auto us = [](auto a, auto b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
        fprintf(stderr, "hp_band_upload_counts: pack+send %ld us, launches %ld us, wait %ld us (%zu bytes over PCIe, %d chunks)\n",
                us(t_begin, t_packed), us(t_packed, t_launched), us(t_launched, t_end), sent, nchunk);
    }
    ctx->n = n; ctx->num = num; ctx->bal_first = bf; ctx->pitch = pitch; ctx->plane = plane;
    ctx->have_band = true;
    return HP_OK;
}

// inspection: what = 0 IR[num], 1 B1[n], 2 balanced band [num][n] (plain layout)
extern "C" int hp_dump_band(hp_ctx* ctx, int32_t what, double* out, int64_t capacity) {
    if (!ctx || !out) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->have_band) return fail(ctx, HP_ERR_STATE, "upload a band first");
    CK(cudaSetDevice(ctx->device));
    if (what == 0 || what == 1) {
        const int64_t cnt = what == 0 ? ctx->num : ctx->n;
        if (capacity < cnt) return fail(ctx, HP_ERR_CAPACITY, "buffer too small");
        CK(cudaMemcpyAsync(out, what == 0 ? ctx->d_ir : ctx->d_b1, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(stream_sync(ctx));
        return HP_OK;
    }
    if (what != 2) return fail(ctx, HP_ERR_INVALID, "bad selector");
    if (capacity < (int64_t)ctx->num * ctx->n) return fail(ctx, HP_ERR_CAPACITY, "buffer too small");
    std::vector<double> tmp(ctx->plane);
    CK(cudaMemcpyAsync(tmp.data(), ctx->d_bal, ctx->plane * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(stream_sync(ctx));
    for (int d = 0; d < ctx->num; ++d)
        for (int64_t r = 0; r < ctx->n; ++r) out[(size_t)d * ctx->n + r] = tmp[qidx(d, (int)r, ctx->pitch)];
    return HP_OK;
}

// ---------------------------------------------------------------------------------------------
// sweep program: callers.py:15-23 (step order) and :138-198 (offsets of one step)
static int build_program(hp_ctx* ctx, const hp_hiccups_params& P) {
    Prog& G = ctx->prog;
    memset(&G, 0, sizeof(G));
    ctx->opa.clear(); ctx->opb.clear(); ctx->opy.clear(); ctx->opr.clear(); ctx->ropi.clear();
    G.npw = P.npw; G.thr = P.min_local_reads;
    int minp = P.pw[0];
    for (int i = 0; i < P.npw; ++i) { G.pw[i] = P.pw[i]; G.ww[i] = P.ww[i]; minp = std::min(minp, P.pw[i]); }
    struct St { int w, p, pi; };
    std::vector<St> steps;
    for (int i = 0; i < P.npw; ++i)
        for (int w = P.ww[i]; w <= P.maxww; ++w) steps.push_back({w, P.pw[i], i});
    std::stable_sort(steps.begin(), steps.end(), [](const St& a, const St& b) { return a.w != b.w ? a.w < b.w : a.p < b.p; });
    if ((int)steps.size() > HP_MAX_STEPS) return fail(ctx, HP_ERR_INVALID, "too many sweep steps");
    if (steps.empty()) return fail(ctx, HP_ERR_INVALID, "no sweep step (ww > maxww)");
    G.nsteps = (int)steps.size();
    bool limit = false;
    int last_p = 0, last_w = 0, nr = 0;
    for (int s = 0; s < G.nsteps; ++s) {
        const int p = steps[s].p, w = steps[s].w;
        for (int a = -w; a <= w; ++a)
            for (int b = -w; b <= w; ++b) {
                const int g = std::max(abs(a), abs(b));
                if (limit && ((g <= last_w && g > std::max(p, last_p)) || g <= std::min(p, last_p))) continue;
                if (a == 0 || b == 0) continue;
                if (abs(a) <= p && abs(b) <= p) continue;
                const bool isy = a > 0 && b < 0;
                const bool isr = isy && (!limit || (p == minp && g > last_w));
                if (isr) ctx->ropi.push_back((unsigned short)ctx->opa.size());
                ctx->opa.push_back((signed char)a); ctx->opb.push_back((signed char)b);
                ctx->opy.push_back(isy); ctx->opr.push_back(isr);
                nr += isr;
            }
        limit = true; last_p = p; last_w = w;
        G.step_pi[s] = steps[s].pi; G.step_w[s] = w;
        G.op_end[s] = (int)ctx->opa.size();
        G.rop_end[s] = nr;
    }
    if ((int)ctx->opa.size() > kMaxOps || nr > kMaxROps)
        return fail(ctx, HP_ERR_INVALID, "sweep program too long for this build (" + std::to_string(ctx->opa.size()) + " offsets)");
    return HP_OK;
}

// host-only inspection hook (no GPU, no context): the sweep program the engine derives from (pw, ww, maxww) -- steps in
// execution order and, per step, the cells it adds in fp64 addition order with their lower-left / Reads flags
extern "C" int hp_program_dump(const hp_hiccups_params* prm, int32_t* nsteps, int32_t* step_p, int32_t* step_w, int32_t* op_end,
                               int64_t op_capacity, int8_t* opa, int8_t* opb, uint8_t* opy, uint8_t* opr, int64_t* nops) {
    if (!prm || !nsteps || !step_p || !step_w || !op_end || !nops) return fail(nullptr, HP_ERR_INVALID, "NULL argument");
    if (prm->npw < 1 || prm->npw > HP_MAX_PW || prm->maxww < 1 || prm->maxww > HP_MAX_WW)
        return fail(nullptr, HP_ERR_INVALID, "npw / maxww out of range");
    hp_ctx tmp;                                        // plain host object: build_program touches no device state
    const int rc = build_program(&tmp, *prm);
    if (rc) return rc;
    *nsteps = tmp.prog.nsteps;
    *nops = (int64_t)tmp.opa.size();
    for (int s = 0; s < tmp.prog.nsteps; ++s) {
        step_p[s] = tmp.prog.pw[tmp.prog.step_pi[s]];
        step_w[s] = tmp.prog.step_w[s];
        op_end[s] = tmp.prog.op_end[s];
    }
    if (opa && opb && opy && opr) {
        if (op_capacity < *nops) return fail(nullptr, HP_ERR_CAPACITY, "op buffers too small");
        for (size_t i = 0; i < tmp.opa.size(); ++i) { opa[i] = tmp.opa[i]; opb[i] = tmp.opb[i]; opy[i] = tmp.opy[i]; opr[i] = tmp.opr[i]; }
    }
    return HP_OK;
}

// quad-interleaved plane: dims (row quad, row & 3, diagonal)
static int make_map_plane(hp_ctx* ctx, CUtensorMap* map, CUtensorMapDataType dt, int esize, void* base, int box_q, int box_d) {
    cuuint64_t dims[3] = {(cuuint64_t)ctx->pitch / 4, 4, (cuuint64_t)ctx->num};
    cuuint64_t strides[2] = {(cuuint64_t)ctx->pitch / 4 * esize, (cuuint64_t)ctx->pitch * esize};
    cuuint32_t box[3] = {(cuuint32_t)box_q, 4, (cuuint32_t)box_d};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ctx->encode(map, dt, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, HP_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return HP_OK;
}

// plain 2-D tensor: dims (inner, outer), outer stride in bytes (a multiple of 16)
static int make_map_2d(hp_ctx* ctx, CUtensorMap* map, CUtensorMapDataType dt, void* base, uint64_t dim0, uint64_t dim1, uint64_t stride_bytes,
                       int box0, int box1) {
    cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ctx->encode(map, dt, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, HP_ERR_CUDA, "cuTensorMapEncodeTiled (2-D) failed: " + std::to_string((int)r));
    return HP_OK;
}

// ---- specialised score kernels: sweep programs compiled in (hp_score_spec.cuh) --------------------
struct SpecKernel {
    bool (*matches)(const Prog&, int, const signed char*, const signed char*, const unsigned char*, const unsigned char*);
    int (*launch)(hp_ctx*, const CUtensorMap&, const ScoreArgs&, dim3, size_t, cudaStream_t);
    int (*launch_levels)(hp_ctx*, const CUtensorMap&, const LevelArgs&, dim3, size_t, cudaStream_t);
    const char* name;
};
template <class PG>
static int launch_levels_spec(hp_ctx* ctx, const CUtensorMap& tm, const LevelArgs& A, dim3 grid, size_t smem, cudaStream_t st) {
    static std::atomic<size_t> granted[64];
    CK(want_smem(k_levels_spec<PG>, ctx->device, smem, granted));
    k_levels_spec<PG><<<grid, kThreads, smem, st>>>(tm, A);
    return HP_OK;
}
template <class PG>
static int launch_spec(hp_ctx* ctx, const CUtensorMap& tm, const ScoreArgs& A, dim3 grid, size_t smem, cudaStream_t st) {
    static std::atomic<size_t> granted[64];
    CK(want_smem(k_score_spec<PG>, ctx->device, smem, granted));
    k_score_spec<PG><<<grid, kSpecThreads, smem, st>>>(tm, A);
    return HP_OK;
}
#define HP_SPEC(name, ...) {spec_matches<SProg<__VA_ARGS__>>, launch_spec<SProg<__VA_ARGS__>>, launch_levels_spec<SProg<__VA_ARGS__>>, name}
static const SpecKernel g_specs[] = {
    HP_SPEC("p2w5", 10, 1, 2, 5),
#ifndef HP_FAST_BUILD       // development builds compile one program only (each takes about a minute)
    HP_SPEC("p1w3", 10, 1, 1, 3),
    HP_SPEC("p4w7", 10, 1, 4, 7),
    HP_SPEC("p124w357", 10, 3, 1, 3, 2, 5, 4, 7),
#endif
};
static const SpecKernel* find_spec(hp_ctx* ctx, int nexec) {
    for (const SpecKernel& k : g_specs)
        if (k.matches(ctx->prog, nexec, ctx->opa.data(), ctx->opb.data(), ctx->opy.data(), ctx->opr.data())) return &k;
    return nullptr;
}

// ---- re-associated score kernels (hp_score_fast.cuh): any single (p, w) program, widths up to FM ----------------
struct FastMaps { CUtensorMap raw, x, lvl; };        // count tile, fp32 balanced tile, level tile
typedef int (*FastLaunch)(hp_ctx*, const FastMaps&, const FastArgs&, int, cudaStream_t);
struct FastKernel {
    int fm;
    FastLaunch launch;              // one (pw, ww) pair, run-time values (pw <= kFMaxPeak, ww >= kFMinWidth)
    FastLaunch launch_gen;          // the general form: one pair of a union program, or any other single pair
    FastLaunch launch_p1w3, launch_p2w5, launch_p4w7;      // the usual pairs, known at compile time
};
template <int FM, bool GEN, int CP = -1, int CW = -1>
static int launch_fast(hp_ctx* ctx, const FastMaps& M, const FastArgs& A, int grid, cudaStream_t st) {
    static std::atomic<size_t> granted[64];
    const size_t smem = FastLayout<FM, FM>::bytes;
    CK(want_smem(k_score_fast<FM, GEN, CP, CW>, ctx->device, smem, granted));
    k_score_fast<FM, GEN, CP, CW><<<grid, kFThreads, smem, st>>>(M.raw, M.x, M.lvl, A);
    return HP_OK;
}
// the fp32 tile: box of fast_px(FM) diagonals x 96 rows of xf[r][d - dlo]; counts and levels: 3-D boxes (16 row quads, 4, 64
// diagonals) of their quad-interleaved planes
static int make_fast_maps(hp_ctx* ctx, int fm, FastMaps* M) {
    int rc = make_map_plane(ctx, &M->raw, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, ctx->d_raw, kFTR / 4, kFTD);
    if (rc) return rc;
    rc = make_map_2d(ctx, &M->x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ctx->d_xf, (uint64_t)(ctx->num - ctx->bal_first), (uint64_t)ctx->n,
                     (uint64_t)ctx->xf_ndp * 4, fast_px(fm), kFXR);
    if (rc) return rc;
    return make_map_plane(ctx, &M->lvl, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, ctx->d_lvl, kFTR / 4, kFTD);
}
#ifdef HP_FAST_BUILD
#define HP_FASTK(FM) {FM, launch_fast<FM, false>, launch_fast<FM, true>, nullptr, launch_fast<FM, false, 2, 5>, nullptr}
#else
#define HP_FASTK(FM) {FM, launch_fast<FM, false>, launch_fast<FM, true>, launch_fast<FM, false, 1, 3>, launch_fast<FM, false, 2, 5>, launch_fast<FM, false, 4, 7>}
#endif
static const FastKernel g_fast[] = {HP_FASTK(8), HP_FASTK(10)};
static FastLaunch pick_fast(const FastKernel* k, int npw, int p, int w) {
    if (npw != 1) return k->launch_gen;
    if (p == 1 && w == 3 && k->launch_p1w3) return k->launch_p1w3;
    if (p == 2 && w == 5 && k->launch_p2w5) return k->launch_p2w5;
    if (p == 4 && w == 7 && k->launch_p4w7) return k->launch_p4w7;
    return (p <= kFMaxPeak && w >= kFMinWidth) ? k->launch : k->launch_gen;
}
static const FastKernel* find_fast(int frozen) {
    for (const FastKernel& k : g_fast)
        if (frozen <= k.fm) return &k;
    return nullptr;
}

static int64_t band_pixel_count(int64_t n, int64_t lo, int64_t hi) {
    hi = std::min(hi, n - 1);
    if (hi < lo) return 0;
    const int64_t k = hi - lo + 1;
    return k * n - (lo + hi) * k / 2;
}

extern "C" int hp_hiccups_score(hp_ctx* ctx, const hp_hiccups_params* prm, hp_hiccups_summary* out) {
    if (!ctx || !prm) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->have_band) return fail(ctx, HP_ERR_STATE, "hp_band_upload must come first");
    const hp_hiccups_params& P = *prm;
    if (P.npw < 1 || P.npw > HP_MAX_PW) return fail(ctx, HP_ERR_INVALID, "npw out of range");
    if (P.maxww < 1 || P.maxww > HP_MAX_WW) return fail(ctx, HP_ERR_INVALID, "maxww out of range [1,20]");
    int minww = P.ww[0], maxw0 = P.ww[0];
    for (int i = 0; i < P.npw; ++i) {
        if (P.pw[i] < 0 || P.ww[i] < 1 || P.ww[i] > P.maxww) return fail(ctx, HP_ERR_INVALID, "need 0 <= pw, 1 <= ww <= maxww");
        for (int j = 0; j < i; ++j) if (P.pw[j] == P.pw[i]) return fail(ctx, HP_ERR_INVALID, "duplicate pw value");
        minww = std::min(minww, P.ww[i]); maxw0 = std::max(maxw0, P.ww[i]);
    }
    if (minww != ctx->bal_first) return fail(ctx, HP_ERR_INVALID, "band was uploaded with bal_first != min(ww)");
    if (!(P.sig >= 0.0)) return fail(ctx, HP_ERR_INVALID, "sig must be >= 0");
    CK(cudaSetDevice(ctx->device));
    ctx->scored = false; ctx->fdr_done = false;
    ctx->prm = P;
    int rc = build_program(ctx, P);
    if (rc) return rc;
    Prog& G = ctx->prog;
    hp_hiccups_summary& S = ctx->sum;
    memset(&S, 0, sizeof(S));
    const int n = (int)ctx->n, num = ctx->num, pitch = ctx->pitch;
    const int dlo = minww;
    const int dhi = (int)std::min<int64_t>(std::min<int64_t>(P.maxapart_bins, num - 1), n - 1);
    ctx->dlo = dlo; ctx->dhi = dhi;
    S.band_pixels = band_pixel_count(n, dlo, std::min<int64_t>(P.maxapart_bins, n - 1));
    if (dhi < dlo) return fail(ctx, HP_ERR_EMPTY_REFIDX, "no band pixel: the reference fails at callers.py:205-208");
    cudaStream_t st = ctx->stream;
    const int nops = (int)ctx->opa.size();
    int launches = 0;

    // ---- tables (program, cell list) -> device ------------------------------------------------------
    const bool force_generic = (P.flags & HP_PF_GENERIC_KERNEL) != 0;
    const bool bhfdr = (P.flags & HP_PF_BHFDR) != 0;
    if (bhfdr && P.npw != 1) return fail(ctx, HP_ERR_INVALID, "the BH-FDR caller takes one (pw, ww) pair");
    const bool spec_ok = !force_generic && num <= 32767;
    Tables& HT = *ctx->h_tab;
    G.nsteps_exec = G.nsteps;
    HT.prog = G;
    memcpy(HT.opa, ctx->opa.data(), nops);
    memcpy(HT.opb, ctx->opb.data(), nops);
    memcpy(HT.opy, ctx->opy.data(), nops);
    memcpy(HT.ropi, ctx->ropi.data(), ctx->ropi.size() * sizeof(unsigned short));
    CK(cudaMemcpyAsync(ctx->d_tab, &HT, sizeof(Tables), cudaMemcpyHostToDevice, st));

    // ---- K1: levels --------------------------------------------------------------------------
    const int F1 = P.maxww;
    CK(cudaMemsetAsync(ctx->d_lhist, 0, (HP_MAX_STEPS + 2) * sizeof(unsigned long long), st));
    {
        const SpecKernel* lspec = spec_ok ? find_spec(ctx, G.nsteps) : nullptr;
        LevelArgs A{};
        A.tab = ctx->d_tab; A.lvl = ctx->d_lvl; A.hist = ctx->d_lhist;
        A.n = n; A.pitch = pitch; A.dlo = dlo; A.dhi = dhi; A.F = F1;
        CUtensorMap tm_raw;
        if (lspec) {
            A.TD = 64; A.BD = A.TD + 3 + 2 * F1; A.NQ = kNQL;
            rc = make_map_plane(ctx, &tm_raw, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, ctx->d_raw, A.NQ, A.BD);
            if (rc) return rc;
            const size_t smem = (size_t)A.BD * 4 * A.NQ * 4 + 16 + (G.nsteps + 2) * 4;
            dim3 grid((n + kTR - 1) / kTR, (dhi - dlo + 3 + A.TD) / A.TD);
            CK(cudaEventRecord(ctx->ev[0], st));      // (right before the launch: host work in between would count as kernel time)
            rc = lspec->launch_levels(ctx, tm_raw, A, grid, smem, st);
            if (rc) return rc;
        } else {
            A.TD = 32; A.BD = A.TD + 2 * F1; A.NQ = ((kTR + F1 + 3) / 4 + 3) & ~3;   // box rows: a multiple of 16 bytes
            rc = make_map_plane(ctx, &tm_raw, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, ctx->d_raw, A.NQ, A.BD);
            if (rc) return rc;
            const size_t smem = (size_t)A.BD * 4 * A.NQ * 4 + 16 + (G.nsteps + 2) * 4;
            static std::atomic<size_t> granted[64];
            CK(want_smem(k_levels, ctx->device, smem, granted));
            dim3 grid((n + kTR - 1) / kTR, (dhi - dlo + A.TD) / A.TD);
            CK(cudaEventRecord(ctx->ev[0], st));
            k_levels<<<grid, kThreads, smem, st>>>(tm_raw, A);
        }
        ++launches;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    unsigned long long* hres_lh = (unsigned long long*)ctx->h_res;                 // [HP_MAX_STEPS + 2]
    CK(cudaMemcpyAsync(hres_lh, ctx->d_lhist, (size_t)(G.nsteps + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(stream_sync(ctx));
    const std::vector<unsigned long long> lh(hres_lh, hres_lh + G.nsteps + 1);

    // ---- replay of the adaptive-width control flow (callers.py:203-232) ------------------------
    unsigned long long total = 0;
    for (auto v : lh) total += v;
    S.n_pixels = (int64_t)total;
    std::vector<unsigned long long> cum(G.nsteps + 1);
    { unsigned long long c = 0; for (int s = 0; s <= G.nsteps; ++s) { c += lh[s]; cum[s] = c; } }
    unsigned long long ini[HP_MAX_PW];
    int lastp[HP_MAX_PW];
    for (int i = 0; i < P.npw; ++i) { ini[i] = total; lastp[i] = -1; }
    int frozen = P.maxww, nexec = 0;
    for (int s = 0; s < G.nsteps; ++s) {
        const int pi = G.step_pi[s], w = G.step_w[s];
        if (w > frozen) break;                      // steps are sorted by w: the rest is skipped too
        if (ini[pi] == 0)
            return fail(ctx, HP_ERR_EMPTY_REFIDX, "unresolved set of p=" + std::to_string(P.pw[pi]) + " is empty at step (" +
                                                       std::to_string(P.pw[pi]) + "," + std::to_string(w) +
                                                       "): the reference fails at callers.py:205-208");
        const unsigned long long resolved = cum[s] - (lastp[pi] >= 0 ? cum[lastp[pi]] : 0ull);
        const double valid = (double)resolved / (double)ini[pi];
        ini[pi] -= resolved;
        const double left = (double)ini[pi] / (double)total;
        S.steps[nexec] = hp_step_stat{P.pw[pi], w, (int64_t)resolved, valid, left};
        lastp[pi] = s;
        ++nexec;
        if (w >= maxw0 && (valid < 0.3 || left < 0.03)) frozen = w;
    }
    S.frozen_w = frozen; S.n_steps = nexec;
    G.nsteps_exec = nexec;
    for (int pi = 0; pi < P.npw; ++pi)
        for (int sv = 0; sv <= G.nsteps + 1; ++sv) {
            unsigned char r = kNoStep;
            for (int s = sv; s < nexec; ++s) if (G.step_pi[s] == pi) { r = (unsigned char)s; break; }
            G.next_step[pi][sv] = r;
        }
    for (int s = 0; s < nexec; ++s) {
        int prev = -1;
        for (int t = 0; t < s; ++t) if (G.step_pi[t] == G.step_pi[s]) prev = t;
        G.step_lo[s] = (unsigned char)(prev + 1);
    }
    G.dspan = maxw0 - minww;
    for (int k = 0; k <= G.dspan; ++k)
        for (int sv = 0; sv <= G.nsteps + 1; ++sv) {
            int last = -1;
            for (int pi = 0; pi < P.npw; ++pi)
                if (minww + k >= P.ww[pi] && G.next_step[pi][sv] != kNoStep) last = std::max(last, (int)G.next_step[pi][sv]);
            G.last_need[k][sv] = last < 0 ? kNoStep : (unsigned char)last;
        }

    // ---- tables for K2 ------------------------------------------------------------------------
    const int F = frozen;
    const int sh_pairs = std::min(P.npw, kShPairs);
    const SpecKernel* spec = spec_ok ? find_spec(ctx, nexec) : nullptr;
    ScoreArgs A{};
    size_t smem = 0;
    dim3 grid;
    if (spec) {
        // 4x4 register-block kernel: one CTA per SM (all 8 warps share one big tile), TD diagonals per CTA
        A.HR = kHR; A.NQ = kNQ;
        int TD = 64;
        auto planes = [&](int td) { return (td + 3 + 4 * F + kTileParts - 1) / kTileParts * kTileParts; };   // whole TMA parts
        while (TD > 8 && score_smem_bytes(planes(TD), kNQ, sh_pairs, kSpecThreads / 32, kQCap, nexec, TD + 8) > 224 * 1024) TD /= 2;
        A.TD = TD; A.BD = planes(TD);
        smem = score_smem_bytes(A.BD, kNQ, sh_pairs, kSpecThreads / 32, kQCap, nexec, TD + 8);
        grid = dim3((n + kTR - 1) / kTR, (dhi - dlo + 3 + TD) / TD);
    } else {
        A.HR = (F + 7) & ~7; A.NQ = (kTR + 2 * A.HR) / 4;
        int TD = 32;
        while (TD > 8 && score_smem_bytes(TD + 4 * F, A.NQ, sh_pairs, 0, 0, 0, 0) > 110 * 1024) TD /= 2;   // two CTAs per SM where possible
        A.TD = TD; A.BD = TD + 4 * F;
        smem = score_smem_bytes(A.BD, A.NQ, sh_pairs, 0, 0, 0, 0);
        grid = dim3((n + kTR - 1) / kTR, (dhi - dlo + TD) / TD);
    }
    if (smem > 227 * 1024) return fail(ctx, HP_ERR_INVALID, "tile does not fit shared memory");
    {
        // candidate thresholds for this sig (host copy of the universal Poisson table)
        Chunks& C = ctx->chunks;
        // kcand[i] = first count k with p(i, k) <= sig.  The table has ~6e5 entries and a linear scan per call reads
        // most of it (0.4 ms of host time per chromosome): the thresholds are kept per (sig, mode) and found by
        // bisection (the tail is decreasing in k), with a walk back over ties / rounding wiggles.
        if (!(ctx->kcand_valid && ctx->kcand_sig == P.sig && ctx->kcand_bhfdr == bhfdr)) {
            const double lim = P.sig * (1.0 + 1e-9) + 1e-300;
            for (int i = 1; i <= C.maxchunk; ++i) {
                const double* p = ctx->h_ptab.data() + C.hoff[i];
                int lo = 0, hi = C.hw[i];                          // answer in [lo, hi]; p[hw - 1] == 0 <= lim
                while (lo < hi) {
                    const int mid = (lo + hi) / 2;
                    if (p[mid] <= lim) hi = mid; else lo = mid + 1;
                }
                while (lo > 0 && p[lo - 1] <= lim) --lo;
                C.kcand[i] = lo;
            }
            if (bhfdr) {   // per-pixel rates: the tail at the LOWER edge of the chunk bounds p from below
                for (int i = C.maxchunk; i >= 2; --i) C.kcand[i] = C.kcand[i - 1];
                C.kcand[1] = 0;
            }
            ctx->kcand_valid = true; ctx->kcand_sig = P.sig; ctx->kcand_bhfdr = bhfdr;
        }
        HT.prog = G;                       // now with the executed steps and the resolve tables
        HT.chunks = C;
        CK(cudaMemcpyAsync(ctx->d_tab, &HT, offsetof(Tables, opa), cudaMemcpyHostToDevice, st));
    }
    const size_t tb = ctx->chunks.total_bins;
    CK(ensure(&ctx->d_betab, &ctx->cap_betab, (size_t)(1 + 2 * F) * 2 * nexec * num));
    CK(ensure(&ctx->d_hist, &ctx->cap_hist, (size_t)P.npw * 2 * tb));
    CK(ensure(&ctx->d_qtab, &ctx->cap_qtab, (size_t)P.npw * 2 * tb));
    if (P.dump) {
        CK(ensure(&ctx->d_dump, &ctx->cap_dump, (size_t)P.npw * 6 * ctx->plane));
        const long long cnt = (long long)P.npw * 6 * ctx->plane;
        k_fill_f64<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(ctx->d_dump, nan(""), cnt);
        ++launches;
    }
    size_t want = std::min<size_t>((size_t)total * P.npw, (size_t)((double)total * P.npw * std::max(0.15, 1.5 * P.sig)) + 65536);
    want = std::max<size_t>(want, 65536);
    // the re-associated kernel: HiCCUPS mode, no per-pixel dump, widths the compiled kernels cover; one launch per pair
    static const bool no_fast_env = getenv("HP_NO_FAST") != nullptr;
    const FastKernel* fast = nullptr;
    bool fast_ok = spec_ok && !no_fast_env && !(P.flags & HP_PF_EXACT_SUMS) && !bhfdr && !P.dump && ctx->domain_ok && ctx->edges_regular;
    for (int i = 0; i < P.npw; ++i) fast_ok = fast_ok && P.pw[i] < P.ww[i];
    if (fast_ok) fast = find_fast(F);
    // per pair: the widths (codes) it resolves at among the executed steps
    int pair_nc[HP_MAX_PW] = {0};
    for (int t = 0; t < nexec; ++t) ++pair_nc[G.step_pi[t]];
    FfsLayout ffsl{};
    size_t want_x = std::max<size_t>(65536, (size_t)total / 8);
    CUtensorMap tm_bal;
    FastMaps fmaps;
    // the specialised kernel loads its tile as kTileParts boxes of consecutive planes
    rc = make_map_plane(ctx, &tm_bal, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, ctx->d_bal, A.NQ,
                        spec ? (A.BD + kTileParts - 1) / kTileParts : A.BD);
    if (rc) return rc;
    if (fast) {
        rc = make_fast_maps(ctx, fast->fm, &fmaps);
        if (rc) return rc;
        CK(ensure(&ctx->d_ffac, &ctx->cap_ffac, (size_t)(1 + 2 * F) * 2 * nexec * num));
        if (!ctx->d_fscratch) CK(cudaMalloc(&ctx->d_fscratch, (size_t)ctx->sm_count * kFScratch * sizeof(int4)));
        const size_t per_code = (size_t)((num - 1 - dlo) / kFTD + 1) * 2 * kFTD;          // k_betab writes every diagonal of the band
        size_t nffs = 0;
        for (int i = 0; i < P.npw; ++i) { ffsl.off[i] = (int)nffs; ffsl.nc[i] = pair_nc[i]; nffs += per_code * pair_nc[i]; }
        CK(ensure(&ctx->d_ffs, &ctx->cap_ffs, nffs));
        CK(cudaMemsetAsync(ctx->d_ffs, 0, nffs * sizeof(float), st));       // diagonals beyond the band: factor 0
        ffsl.base = ctx->d_ffs;
    }
    k_betab<<<dim3((num + kBetabThreads - 1) / kBetabThreads, 1 + 2 * F), kBetabThreads, 0, st>>>(ctx->d_tab, ctx->d_ir, ctx->d_betab, num, ctx->bal_first, nexec, F,
                                                                     fast ? ctx->d_ffac : nullptr, ffsl);
    ++launches;
    unsigned int cnt[16] = {0};
    unsigned long long small[48];
    bool use_fast = fast != nullptr;
    for (int attempt = 0;; ++attempt) {
        CK(ensure(&ctx->d_cand, &ctx->cap_cand, use_fast ? std::max<size_t>(65536, want_x) : want));
        if (use_fast) {
            CK(ensure(&ctx->d_fcand, &ctx->cap_fcand, want + (size_t)kFCandChunk * kFWarps * ctx->sm_count * P.npw));   // + every warp's last piece, per launch
            CK(ensure(&ctx->d_xrec, &ctx->cap_xrec, want_x));
        }
        CK(cudaMemsetAsync(ctx->d_hist, 0, (size_t)P.npw * 2 * tb * sizeof(unsigned int), st));
        CK(cudaMemsetAsync(ctx->d_small, 0, 48 * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(ctx->d_cnt, 0, 48 * sizeof(unsigned int), st));     // [16..31]: the fast kernel's running E.max() bounds per pair
        A.tab = ctx->d_tab; A.raw = ctx->d_raw; A.lvl = ctx->d_lvl; A.ir = ctx->d_ir; A.b1 = ctx->d_b1; A.b2 = ctx->d_b2;
        A.betab = ctx->d_betab; A.hist = ctx->d_hist; A.emax_bits = ctx->d_small; A.nvalid = ctx->d_small + 16;
        A.cand = ctx->d_cand; A.cand_count = ctx->d_cnt;
        A.cand_cap = (unsigned)std::min<size_t>(ctx->cap_cand, 0xffffffffu);
        A.dump = P.dump ? ctx->d_dump : nullptr; A.plane = (long long)ctx->plane;
        A.n = n; A.num = num; A.pitch = pitch; A.dlo = dlo; A.dhi = dhi; A.F = F;
        A.bal_first = ctx->bal_first; A.sh_pairs = sh_pairs;
        A.bhfdr = bhfdr ? 1 : 0;
        A.nexec = nexec; A.npw = P.npw; A.dspan = G.dspan; A.maxchunk = ctx->chunks.maxchunk; A.total_bins = ctx->chunks.total_bins;
        for (int i = 0; i < P.npw; ++i) A.ww[i] = P.ww[i];
        for (int k = 0; k < nexec; ++k) { A.step_pi[k] = (unsigned char)G.step_pi[k]; A.step_lo[k] = G.step_lo[k]; }
        memcpy(A.last_need, G.last_need, sizeof(A.last_need));
        if (!use_fast) CK(cudaEventRecord(ctx->ev[2], st));      // ms_score = the score kernel alone (roofline leg of bench.py)
        if (use_fast) {
            FastArgs FA{};
            FA.ffac = ctx->d_ffac; FA.ffs = ctx->d_ffs; FA.b1f = ctx->d_b1f; FA.b2s = ctx->d_b2s;
            FA.tab = ctx->d_tab; FA.scratch = ctx->d_fscratch;
            FA.fcand = ctx->d_fcand; FA.xrec = ctx->d_xrec; FA.cnt = ctx->d_cnt;
            FA.fcand_cap = (unsigned)std::min<size_t>(ctx->cap_fcand, 0xffffffffu);
            FA.xrec_cap = (unsigned)std::min<size_t>(ctx->cap_xrec, 0xffffffffu);
            FA.n = n; FA.num = num; FA.pitch = pitch; FA.dlo = dlo; FA.dhi = dhi; FA.F = F; FA.nexec = nexec;
            FA.maxchunk = ctx->chunks.maxchunk; FA.total_bins = ctx->chunks.total_bins;
            FA.nstrips = (dhi - dlo) / kFTD + 1; FA.ntr = (n + kFTR - 1) / kFTR;
            {   // mantissa bits of the in-octave edges, rounded down / up with margin (fast_classify)
                auto mant = [](double x, bool up) {
                    float f = (float)(x * (up ? 1.0 + 1e-9 : 1.0 - 1e-9));
                    f = nextafterf(f, up ? 4.0f : 0.0f);
                    unsigned u; memcpy(&u, &f, 4);
                    return u & 0x7fffffu;
                };
                FA.c1dn = mant(ctx->chunks.rv[2], false); FA.c2dn = mant(ctx->chunks.rv[3], false);
                FA.c1up = mant(ctx->chunks.rv[2], true); FA.c2up = mant(ctx->chunks.rv[3], true);
                // lower edge of every chunk (rv[i - 1]), rounded up with the same margin; chunk maxchunk + 1 = beyond the last edge
                for (int i = 0; i < kMaxChunk + 4; ++i) FA.elo[i] = INFINITY;
                FA.elo[0] = 0.f; FA.elo[1] = 0.f;
                for (int i = 2; i <= ctx->chunks.maxchunk + 1; ++i)
                    FA.elo[i] = nextafterf((float)(ctx->chunks.rv[i - 1] * (1.0 + 1e-9)), INFINITY);
            }
            const int items = FA.nstrips * FA.ntr;
            // every launch's arguments first, then the event and the launches back to back: whatever the host does between
            // the event and a launch would be counted as kernel time (with 8 ranks on one box the host is the slow side)
            std::vector<FastArgs> fas;
            std::vector<FastLaunch> fns;
            for (int pi = 0; pi < P.npw; ++pi) {
                if (pair_nc[pi] == 0) continue;                // no executed step of this pair: nothing resolves for it
                FA.pair = pi; FA.p = P.pw[pi]; FA.w0 = P.ww[pi]; FA.wpair = P.ww[pi]; FA.ncode = pair_nc[pi];
                FA.ffs = ctx->d_ffs + ffsl.off[pi];
                FA.hist = ctx->d_hist + (size_t)pi * 2 * tb; FA.nvalid = ctx->d_small + 16 + 2 * pi; FA.gmax = ctx->d_cnt + 16 + 2 * pi;
                FA.nlevels = G.nsteps;
                memset(FA.tcode, 0xF, sizeof(FA.tcode)); memset(FA.tstep, 0, sizeof(FA.tstep));
                memset(FA.hmask, 0, sizeof(FA.hmask)); memset(FA.cabs, 0, sizeof(FA.cabs)); memset(FA.ctab, 0, sizeof(FA.ctab));
                if (pair_nc[pi] > kFMaxCode - 1) return fail(ctx, HP_ERR_INVALID, "internal: too many widths for the fast kernel");
                for (int lv = 0; lv < G.nsteps; ++lv) {
                    const int t = G.next_step[pi][lv];
                    if (t != kNoStep) FA.tcode[lv] = (unsigned char)(G.step_w[t] - P.ww[pi]);
                }
                for (int t = 0; t < nexec; ++t) {
                    if (G.step_pi[t] != pi) continue;
                    const int code = G.step_w[t] - P.ww[pi];
                    FA.tstep[code] = (unsigned char)t;
                    // ring multiplicities of the accumulators at step t, from the cell list itself
                    int cells[kFMaxG + 2] = {0};
                    for (int k = 0; k < G.op_end[t]; ++k) {
                        const int g = std::max(abs((int)ctx->opa[k]), abs((int)ctx->opb[k]));
                        if (g > F) return fail(ctx, HP_ERR_INVALID, "internal: cell beyond the frozen width");
                        ++cells[g];
                    }
                    int mult[kFMaxG + 2] = {0};
                    for (int g = 1; g <= F; ++g) {
                        if (cells[g] % (8 * g - 4)) return fail(ctx, HP_ERR_INVALID, "internal: partial ring in the sweep program");
                        mult[g] = cells[g] / (8 * g - 4);
                    }
                    for (int g = 1; g <= F; ++g) {
                        const int c = mult[g] - mult[g + 1];
                        FA.ctab[code][g] = (float)c;
                        if (c) FA.hmask[code] |= (unsigned short)(1u << g);
                        FA.cabs[g] = std::max(FA.cabs[g], (float)abs(c));
                    }
                }
                fas.push_back(FA);
                fns.push_back(pick_fast(fast, P.npw, P.pw[pi], P.ww[pi]));
            }
            CK(cudaEventRecord(ctx->ev[2], st));
            for (size_t k = 0; k < fas.size(); ++k) {
                rc = fns[k](ctx, fmaps, fas[k], std::min(items, ctx->sm_count), st);
                if (rc) return rc;
                ++launches;
            }
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev[3], st));
            // the records the fast kernel could not settle + the E.max() contenders, in the reference's fp64 order
            static std::atomic<size_t> granted_x[64];
            const size_t smem_x = ((score_smem_bytes(0, 0, sh_pairs, 0, 0, 0, 0) + 127) & ~(size_t)127) +
                                  (size_t)(kExThreads / 32) * ex_warp_bytes(kExRecFew);
            CK(want_smem(k_exact, ctx->device, smem_x, granted_x));
            k_exact<<<2 * ctx->sm_count, kExThreads, smem_x, st>>>(A, ctx->d_bal, ctx->d_xrec, ctx->d_cnt + 10, FA.xrec_cap);
            ++launches;
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev[6], st));
        } else {
            if (spec) {
                rc = spec->launch(ctx, tm_bal, A, grid, smem, st);
                if (rc) return rc;
            } else {
                static std::atomic<size_t> granted[64];
                CK(want_smem(k_score, ctx->device, smem, granted));
                k_score<<<grid, kThreads, smem, st>>>(tm_bal, A);
            }
            ++launches;
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev[3], st));
        }
        CK(cudaMemcpyAsync(ctx->h_res + 2048, ctx->d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_res + 2048 + 64, ctx->d_small, sizeof(small), cudaMemcpyDeviceToHost, st));
        CK(stream_sync(ctx));
        memcpy(cnt, ctx->h_res + 2048, sizeof(cnt));
        memcpy(small, ctx->h_res + 2048 + 64, sizeof(small));
        if (attempt >= 4) return fail(ctx, HP_ERR_CAPACITY, "candidate buffer overflow");
        if (use_fast) {
            if (cnt[11]) { want_x = (size_t)total * P.npw + 1024; continue; }
            if (cnt[9]) { want = (size_t)total * P.npw + 1024; continue; }
        }
        if (cnt[1] == 0) break;
        want = (size_t)total * P.npw + 1024;      // every pixel can be a candidate at most once per pair
        want_x = std::max(want_x, want);
    }
    ctx->fast_used = use_fast;
    ctx->nfcand = use_fast ? (unsigned)std::min<size_t>(cnt[8], ctx->cap_fcand) : 0u;      // slots handed out (unused ones are marked)
    ctx->spec_used = spec != nullptr || use_fast;
    if (cnt[2]) return fail(ctx, HP_ERR_CHUNK_OVERFLOW,
                            std::to_string(cnt[2]) + " expected values exceed the last lambda-chunk edge " +
                                std::to_string(ctx->chunks.rv[ctx->chunks.maxchunk]) + "; create the context with a larger max_chunks");
    ctx->ncand = cnt[0];
    S.n_candidates = (int64_t)cnt[0] + ctx->nfcand;
    S.fast_kernel = use_fast ? 1 : 0;
    S.n_exact = use_fast ? (int64_t)cnt[10] : 0;
    for (int i = 0; i < P.npw; ++i)
        for (int fl = 0; fl < 2; ++fl) {
            hp_lf_stat& L = S.lf[i][fl];
            L.n_valid = (int64_t)small[16 + i * 2 + fl];
            double em;
            memcpy(&em, &small[i * 2 + fl], 8);
            L.e_max = em;
            L.numbin = (L.n_valid > 0) ? (int)ceil(log(em) / log(2.0) * 3 + 1) : 0;
        }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); S.ms_levels = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); S.ms_score = ms;
    S.ms_exact = 0.f;
    if (use_fast) { cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[6]); S.ms_exact = ms; }
    S.ms_total = S.ms_levels + S.ms_score + S.ms_exact;
    S.launches = launches;
    S.spec_kernel = ctx->spec_used ? 1 : 0;
    ctx->scored = true;
    if (out) *out = S;
    return HP_OK;
}

extern "C" int hp_hiccups_fdr(hp_ctx* ctx, const int32_t* numbin_override, hp_hiccups_summary* out) {
    if (!ctx) return fail(ctx, HP_ERR_INVALID, "NULL ctx");
    if (!ctx->scored) return fail(ctx, HP_ERR_STATE, "hp_hiccups_score must come first");
    CK(cudaSetDevice(ctx->device));
    const hp_hiccups_params& P = ctx->prm;
    hp_hiccups_summary& S = ctx->sum;
    cudaStream_t st = ctx->stream;
    int nb[16] = {0};
    int maxnb = 0;
    for (int i = 0; i < P.npw; ++i)
        for (int fl = 0; fl < 2; ++fl) {
            int v = numbin_override ? numbin_override[i * 2 + fl] : S.lf[i][fl].numbin;
            S.lf[i][fl].numbin = v;
            v = std::max(0, std::min(v, ctx->chunks.maxchunk));
            nb[i * 2 + fl] = v;
            maxnb = std::max(maxnb, v);
        }
    CK(cudaEventRecord(ctx->ev[4], st));
    CK(cudaMemcpyAsync(ctx->d_numbin, nb, sizeof(nb), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_small + 32, 0, 16 * sizeof(unsigned long long), st));
    int launches = 0;
    if (maxnb > 0) {
        CK(cudaMemsetAsync(ctx->d_kq, 0x7f, 16 * (kMaxChunk + 2) * sizeof(int), st));
        k_bh<<<dim3(maxnb, P.npw * 2), kThreads, 0, st>>>(ctx->d_tab, ctx->d_hist, ctx->d_ptab, ctx->d_qtab, ctx->d_numbin, ctx->d_kq, P.sig);
        ++launches;
        CK(cudaGetLastError());
    }
    unsigned int cnt[4] = {0, 0, 0, 0};
    unsigned long long nrej[16] = {0};
    const size_t ncand_all = (size_t)ctx->ncand + ctx->nfcand;
    size_t want = std::max<size_t>(65536, ncand_all / 4 + 1024);
    for (int attempt = 0; attempt < 2 && ncand_all > 0; ++attempt) {
        CK(ensure(&ctx->d_surv, &ctx->cap_surv, want));
        CK(cudaMemsetAsync(ctx->d_cnt + 4, 0, 4 * sizeof(unsigned int), st));
        CK(cudaMemsetAsync(ctx->d_small + 32, 0, 16 * sizeof(unsigned long long), st));
        FilterArgs A{};
        A.tab = ctx->d_tab; A.cand = ctx->d_cand; A.ncand = ctx->ncand; A.ptab = ctx->d_ptab; A.qtab = ctx->d_qtab; A.numbin = ctx->d_numbin;
        A.bal = ctx->d_bal; A.out = ctx->d_surv; A.out_count = ctx->d_cnt + 4;
        A.out_cap = (unsigned)std::min<size_t>(ctx->cap_surv, 0xffffffffu);
        A.nreject = ctx->d_small + 32; A.sig = P.sig; A.pitch = ctx->pitch;
        A.bhfdr = (P.flags & HP_PF_BHFDR) ? 1 : 0;
        if (ctx->ncand) {
            k_filter<<<(ctx->ncand + 255) / 256, 256, 0, st>>>(A);
            ++launches;
            CK(cudaGetLastError());
        }
        if (ctx->nfcand) {
            // the re-associated kernel's candidates: same selection, then the survivors' E in the reference's fp64 order
            const unsigned fblocks = std::min<unsigned>((ctx->nfcand + kFiltThreads - 1) / kFiltThreads, 8u * (unsigned)ctx->sm_count);
            k_filter_fast<<<fblocks, kFiltThreads, 0, st>>>(A, ctx->d_fcand, ctx->nfcand, ctx->d_kq, P.npw);
            ++launches;
            CK(cudaGetLastError());
            FillArgs FA{};
            FA.tab = ctx->d_tab; FA.bal = ctx->d_bal; FA.ir = ctx->d_ir; FA.b1 = ctx->d_b1; FA.b2 = ctx->d_b2; FA.betab = ctx->d_betab;
            FA.surv = ctx->d_surv; FA.nsurv_ptr = ctx->d_cnt + 4; FA.cap = A.out_cap;
            FA.n = (int)ctx->n; FA.num = ctx->num; FA.pitch = ctx->pitch; FA.bal_first = ctx->bal_first; FA.F = S.frozen_w; FA.nexec = S.n_steps;
            static std::atomic<size_t> granted_f[64];
            const size_t smem_f = (size_t)(kFillThreads / 32) * ex_warp_bytes(kExRec);
            CK(want_smem(k_fill_exact, ctx->device, smem_f, granted_f));
            k_fill_exact<<<3 * ctx->sm_count, kFillThreads, smem_f, st>>>(FA);
            ++launches;
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(ctx->h_res + 4096, ctx->d_cnt + 4, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_res + 4096 + 64, ctx->d_small + 32, sizeof(nrej), cudaMemcpyDeviceToHost, st));
        if (attempt == 0) CK(cudaEventRecord(ctx->ev[5], st));
        CK(stream_sync(ctx));
        memcpy(cnt, ctx->h_res + 4096, sizeof(cnt));
        memcpy(nrej, ctx->h_res + 4096 + 64, sizeof(nrej));
        if (cnt[1] == 0) break;
        if (attempt == 1) return fail(ctx, HP_ERR_CAPACITY, "survivor buffer overflow");
        want = ncand_all + 16;
    }
    if (ncand_all == 0) {
        CK(cudaEventRecord(ctx->ev[5], st));
        CK(stream_sync(ctx));
    }
    for (int i = 0; i < P.npw; ++i)
        for (int fl = 0; fl < 2; ++fl) S.lf[i][fl].n_reject = (int64_t)nrej[i * 2 + fl];
    ctx->nsurv = cnt[0];
    S.n_survivors = cnt[0];
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    S.ms_fdr = ms;
    S.ms_total = S.ms_levels + S.ms_score + S.ms_fdr;
    S.launches += launches;
    memcpy(ctx->numbin, nb, sizeof(nb));
    ctx->fdr_done = true;
    if (out) *out = S;
    return HP_OK;
}

extern "C" int hp_get_summary(hp_ctx* ctx, hp_hiccups_summary* out) {
    if (!ctx || !out) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->scored) return fail(ctx, HP_ERR_STATE, "hp_hiccups_score must come first");
    *out = ctx->sum;
    return HP_OK;
}

extern "C" int hp_hiccups(hp_ctx* ctx, const hp_hiccups_params* prm, hp_hiccups_summary* out) {
    int rc = hp_hiccups_score(ctx, prm, nullptr);
    if (rc) return rc;
    return hp_hiccups_fdr(ctx, nullptr, out);
}

extern "C" int hp_get_survivors(hp_ctx* ctx, hp_survivor* buf, int64_t capacity, int64_t* count) {
    if (!ctx || !count) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->fdr_done) return fail(ctx, HP_ERR_STATE, "hp_hiccups_fdr must come first");
    *count = ctx->nsurv;
    if (!buf) return HP_OK;
    if (capacity < (int64_t)ctx->nsurv) return fail(ctx, HP_ERR_CAPACITY, "survivor buffer too small");
    CK(cudaSetDevice(ctx->device));
    if (ctx->nsurv) {
        const size_t bytes = (size_t)ctx->nsurv * sizeof(hp_survivor);
        CK(ensure_host(&ctx->h_out, &ctx->cap_out, bytes));
        CK(cudaMemcpyAsync(ctx->h_out, ctx->d_surv, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(stream_sync(ctx));
        memcpy(buf, ctx->h_out, bytes);
    }
    return HP_OK;
}

extern "C" int hp_hist_bins(hp_ctx* ctx, int64_t* total_bins) {
    if (!ctx || !total_bins) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    *total_bins = ctx->chunks.total_bins;
    return HP_OK;
}

extern "C" int hp_hist_export(hp_ctx* ctx, int64_t* out, int64_t capacity) {
    if (!ctx || !out) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->scored) return fail(ctx, HP_ERR_STATE, "hp_hiccups_score must come first");
    const size_t cnt = (size_t)ctx->prm.npw * 2 * ctx->chunks.total_bins;
    if ((size_t)capacity < cnt) return fail(ctx, HP_ERR_CAPACITY, "histogram buffer too small");
    CK(cudaSetDevice(ctx->device));
    std::vector<unsigned int> h(cnt);
    CK(cudaMemcpyAsync(h.data(), ctx->d_hist, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(stream_sync(ctx));
    for (size_t i = 0; i < cnt; ++i) out[i] = h[i];
    return HP_OK;
}

extern "C" int hp_hist_import(hp_ctx* ctx, const int64_t* in, int64_t count) {
    if (!ctx || !in) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->scored) return fail(ctx, HP_ERR_STATE, "hp_hiccups_score must come first");
    const size_t cnt = (size_t)ctx->prm.npw * 2 * ctx->chunks.total_bins;
    if ((size_t)count != cnt) return fail(ctx, HP_ERR_INVALID, "histogram size mismatch");
    std::vector<unsigned int> h(cnt);
    for (size_t i = 0; i < cnt; ++i) {
        if (in[i] < 0 || in[i] > 0xffffffffll) return fail(ctx, HP_ERR_INVALID, "histogram count out of range");
        h[i] = (unsigned int)in[i];
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_hist, h.data(), cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(stream_sync(ctx));
    ctx->fdr_done = false;
    return HP_OK;
}

extern "C" int hp_get_gaps(hp_ctx* ctx, uint8_t* out, int64_t n) {
    if (!ctx || !out) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->have_band) return fail(ctx, HP_ERR_STATE, "hp_band_upload must come first");
    if (n < ctx->n) return fail(ctx, HP_ERR_CAPACITY, "gap buffer too small");
    CK(cudaSetDevice(ctx->device));
    CK(ensure_host(&ctx->h_out, &ctx->cap_out, (size_t)ctx->n * 4));
    const unsigned int* nz = (const unsigned int*)ctx->h_out;
    CK(cudaMemcpyAsync(ctx->h_out, ctx->d_rownz, (size_t)ctx->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(stream_sync(ctx));
    for (int64_t i = 0; i < ctx->n; ++i) out[i] = nz[i] ? 0 : 1;
    return HP_OK;
}

extern "C" int hp_dump_levels(hp_ctx* ctx, uint8_t* out, int64_t capacity) {
    if (!ctx || !out) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->scored) return fail(ctx, HP_ERR_STATE, "hp_hiccups_score must come first");
    if (capacity < (int64_t)ctx->num * ctx->n) return fail(ctx, HP_ERR_CAPACITY, "level buffer too small");
    CK(cudaSetDevice(ctx->device));
    memset(out, kLvlNone, (size_t)ctx->num * ctx->n);
    const int rows = ctx->dhi - ctx->dlo + 1, pitch = ctx->pitch;
    std::vector<unsigned char> tmp((size_t)rows * pitch);
    CK(cudaMemcpyAsync(tmp.data(), ctx->d_lvl + (size_t)ctx->dlo * pitch, tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(stream_sync(ctx));
    for (int k = 0; k < rows; ++k) {
        uint8_t* o = out + (size_t)(ctx->dlo + k) * ctx->n;
        for (int64_t r = 0; r < ctx->n; ++r) o[r] = tmp[qidx(k, (int)r, pitch)];
    }
    return HP_OK;
}

extern "C" int hp_dump_plane(hp_ctx* ctx, int32_t pair, int32_t background, int32_t what, double* out, int64_t capacity) {
    if (!ctx || !out) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->scored || !ctx->prm.dump || !ctx->d_dump) return fail(ctx, HP_ERR_STATE, "score with params.dump != 0 first");
    if (pair < 0 || pair >= ctx->prm.npw || background < 0 || background > 1 || what < 0 || what > 2)
        return fail(ctx, HP_ERR_INVALID, "bad plane selector");
    if (capacity < (int64_t)ctx->num * ctx->n) return fail(ctx, HP_ERR_CAPACITY, "plane buffer too small");
    CK(cudaSetDevice(ctx->device));
    const double* src = ctx->d_dump + (size_t)((pair * 2 + background) * 3 + what) * ctx->plane;
    CK(cudaMemcpy2DAsync(out, ctx->n * 8, src, (size_t)ctx->pitch * 8, ctx->n * 8, ctx->num, cudaMemcpyDeviceToHost, ctx->stream));
    CK(stream_sync(ctx));
    return HP_OK;
}

extern "C" int hp_get_chunk_table(hp_ctx* ctx, int32_t pair, int32_t background, int32_t* numbin, int32_t* widths,
                                  int64_t* hist, double* p, double* q, int64_t capacity) {
    if (!ctx || !numbin) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (!ctx->fdr_done) return fail(ctx, HP_ERR_STATE, "hp_hiccups_fdr must come first");
    if (pair < 0 || pair >= ctx->prm.npw || background < 0 || background > 1) return fail(ctx, HP_ERR_INVALID, "bad selector");
    const int lf = pair * 2 + background;
    const int nb = ctx->numbin[lf];
    *numbin = nb;
    const Chunks& C = ctx->chunks;
    if (widths) for (int i = 1; i <= nb; ++i) widths[i - 1] = C.hw[i];
    if (!hist && !p && !q) return HP_OK;
    const size_t cnt = nb ? (size_t)C.hoff[nb] + C.hw[nb] : 0;
    if ((size_t)capacity < cnt) return fail(ctx, HP_ERR_CAPACITY, "table buffer too small");
    CK(cudaSetDevice(ctx->device));
    if (cnt == 0) return HP_OK;
    if (hist) {
        std::vector<unsigned int> h(cnt);
        CK(cudaMemcpyAsync(h.data(), ctx->d_hist + (size_t)lf * C.total_bins, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(stream_sync(ctx));
        for (size_t i = 0; i < cnt; ++i) hist[i] = h[i];
    }
    if (p) memcpy(p, ctx->h_ptab.data(), cnt * 8);
    if (q) {
        CK(cudaMemcpyAsync(q, ctx->d_qtab + (size_t)lf * C.total_bins, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(stream_sync(ctx));
    }
    return HP_OK;
}

extern "C" int hp_poisson_sf(hp_ctx* ctx, const double* k, const double* mu, double* out, int64_t count) {
    if (!ctx || !k || !mu || !out || count < 0) return fail(ctx, HP_ERR_INVALID, "bad argument");
    if (count == 0) return HP_OK;
    CK(cudaSetDevice(ctx->device));
    double* d = nullptr;
    CK(cudaMalloc(&d, (size_t)count * 24));
    cudaError_t e = cudaMemcpyAsync(d, k, count * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + count, mu, count * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_poisson_sf<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d, d + count, d + 2 * count, count);
        e = cudaMemcpyAsync(out, d + 2 * count, count * 8, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = stream_sync(ctx);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, HP_ERR_CUDA, cudaGetErrorString(e));
    return HP_OK;
}


// ---------------------------------------------------------------------------------------------
// APA -- /root/reference/hicpeaks/apa.py
static int apa_rec(ApaPlan& P, int start, int n) {
    if (n <= 128) {
        const int l = P.nleaf++;
        P.leaf_start[l] = start; P.leaf_len[l] = n;
        return l;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int a = apa_rec(P, start, n2), b = apa_rec(P, start + n2, n - n2);
    P.comb_dst[P.ncomb] = (short)a; P.comb_src[P.ncomb] = (short)b; ++P.ncomb;
    return a;
}

extern "C" int hp_apa_upload(hp_ctx* ctx, const hp_apa_desc* b) {
    if (!ctx || !b || !b->bal_diags) return fail(ctx, HP_ERR_INVALID, "NULL argument");
    if (b->n <= 0 || b->n > (1ll << 30) || b->num <= 0 || b->num > b->n) return fail(ctx, HP_ERR_INVALID, "bad band geometry");
    CK(cudaSetDevice(ctx->device));
    const int64_t n = b->n;
    const int num = b->num;
    const size_t cells = (size_t)n * num;
    CK(ensure(&ctx->d_apa_bal, &ctx->cap_apa_bal, cells));
    if (cells * 8 > ctx->cap_stage) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->cap_stage = 0;
        CK(cudaHostAlloc(&ctx->h_stage, cells * 8, cudaHostAllocDefault));
        ctx->cap_stage = cells * 8;
    }
    double* h = (double*)ctx->h_stage;                 // diagonal-major staging [d][n]; transposed on the device
    for (int d = 0; d < num; ++d) {
        memcpy(h + (size_t)d * n, b->bal_diags[d], (size_t)(n - d) * 8);
        if (d) memset(h + (size_t)d * n + (n - d), 0, (size_t)d * 8);
    }
    double* tmp = nullptr;
    CK(cudaMalloc(&tmp, cells * 8));
    cudaError_t e = cudaMemcpyAsync(tmp, h, cells * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_apa_transpose<<<dim3((unsigned)((n + 31) / 32), (num + 31) / 32), dim3(32, 8), 0, ctx->stream>>>(tmp, ctx->d_apa_bal, n, num);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = stream_sync(ctx);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(ctx, HP_ERR_CUDA, std::string("APA band upload: ") + cudaGetErrorString(e));
    ctx->apa_n = n; ctx->apa_num = num; ctx->apa_npos = 0;
    return HP_OK;
}

static int apa_plan(hp_ctx* ctx, int cells);
extern "C" int hp_apa_windows(hp_ctx* ctx, const int32_t* pos_i, const int32_t* pos_j, int64_t npos, int32_t w, uint8_t* valid,
                              double* mean_arr) {
    if (!ctx || !pos_i || !pos_j || !valid || !mean_arr || npos < 0) return fail(ctx, HP_ERR_INVALID, "bad argument");
    if (!ctx->apa_num) return fail(ctx, HP_ERR_STATE, "hp_apa_upload must come first");
    const int side = 2 * w + 1, cells = side * side;
    if (w < 0 || cells > 128 * kApaMaxLeaves / 2) return fail(ctx, HP_ERR_INVALID, "window too large");
    for (int64_t k = 0; k < npos; ++k) {
        const int64_t i = pos_i[k], j = pos_j[k];
        const bool inside = i - w >= 0 && i + w + 1 <= ctx->apa_n && j - w >= 0 && j + w + 1 <= ctx->apa_n;
        if (inside && llabs(j - i) + 2 * w >= ctx->apa_num)
            return fail(ctx, HP_ERR_INVALID, "anchor " + std::to_string(k) + " needs diagonals beyond the uploaded band");
    }
    ctx->apa_npos = 0;
    if (npos == 0) { ctx->apa_w = w; return HP_OK; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rcp = apa_plan(ctx, cells);
    if (rcp) return rcp;
    CK(ensure(&ctx->d_apa_pos, &ctx->cap_apa_pos, (size_t)2 * npos));
    CK(ensure(&ctx->d_apa_wins, &ctx->cap_apa_wins, (size_t)npos * cells));
    CK(ensure(&ctx->d_apa_valid, &ctx->cap_apa_valid, (size_t)npos));
    CK(ensure(&ctx->d_apa_mean, &ctx->cap_apa_mean, (size_t)npos));
    CK(cudaMemcpyAsync(ctx->d_apa_pos, pos_i, npos * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_apa_pos + npos, pos_j, npos * 4, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)cells * 8 + (size_t)kApaMaxLeaves * 9 * 8;
    CK(cudaFuncSetAttribute(k_apa_windows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_apa_windows<<<(unsigned)npos, kApaThreads, smem, st>>>(ctx->d_apa_plan, ctx->d_apa_bal, ctx->apa_n, ctx->apa_num, ctx->d_apa_pos,
                                                             ctx->d_apa_pos + npos, w, ctx->d_apa_wins, ctx->d_apa_valid, ctx->d_apa_mean);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(valid, ctx->d_apa_valid, npos, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(mean_arr, ctx->d_apa_mean, npos * 8, cudaMemcpyDeviceToHost, st));
    CK(stream_sync(ctx));
    ctx->apa_npos = npos; ctx->apa_w = w;
    return HP_OK;
}

static int apa_plan(hp_ctx* ctx, int cells) {
    if (!ctx->d_apa_plan) CK(cudaMalloc(&ctx->d_apa_plan, sizeof(ApaPlan)));
    ApaPlan P{};
    P.n = cells;
    apa_rec(P, 0, cells);
    CK(cudaMemcpyAsync(ctx->d_apa_plan, &P, sizeof(P), cudaMemcpyHostToDevice, ctx->stream));
    CK(stream_sync(ctx));      // P lives on this stack frame
    return HP_OK;
}

extern "C" int hp_apa_load_windows(hp_ctx* ctx, const double* wins, int64_t npos, int32_t w, double* mean_arr) {
    if (!ctx || !wins || !mean_arr || npos <= 0) return fail(ctx, HP_ERR_INVALID, "bad argument");
    const int side = 2 * w + 1, cells = side * side;
    if (w < 0 || cells > 128 * kApaMaxLeaves / 2) return fail(ctx, HP_ERR_INVALID, "window too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->apa_npos = 0;
    int rc = apa_plan(ctx, cells);
    if (rc) return rc;
    CK(ensure(&ctx->d_apa_wins, &ctx->cap_apa_wins, (size_t)npos * cells));
    CK(ensure(&ctx->d_apa_mean, &ctx->cap_apa_mean, (size_t)npos));
    CK(cudaMemcpyAsync(ctx->d_apa_wins, wins, (size_t)npos * cells * 8, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)cells * 8 + (size_t)kApaMaxLeaves * 9 * 8;
    CK(cudaFuncSetAttribute(k_apa_means, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_apa_means<<<(unsigned)npos, kApaThreads, smem, st>>>(ctx->d_apa_plan, ctx->d_apa_wins, cells, ctx->d_apa_mean);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(mean_arr, ctx->d_apa_mean, npos * 8, cudaMemcpyDeviceToHost, st));
    CK(stream_sync(ctx));
    ctx->apa_npos = npos; ctx->apa_w = w;
    return HP_OK;
}

extern "C" int hp_apa_accumulate(hp_ctx* ctx, const int64_t* sel, int64_t nsel, double* acc, int32_t init) {
    if (!ctx || !acc || nsel < 0 || (nsel && !sel)) return fail(ctx, HP_ERR_INVALID, "bad argument");
    if (!ctx->apa_npos) return fail(ctx, HP_ERR_STATE, "hp_apa_windows must come first");
    for (int64_t k = 0; k < nsel; ++k)
        if (sel[k] < 0 || sel[k] >= ctx->apa_npos) return fail(ctx, HP_ERR_INVALID, "window index out of range");
    if (nsel == 0) return HP_OK;
    const int side = 2 * ctx->apa_w + 1, cells = side * side;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ensure(&ctx->d_apa_sel, &ctx->cap_apa_sel, (size_t)nsel));
    CK(ensure(&ctx->d_apa_avg, &ctx->cap_apa_avg, (size_t)cells));
    CK(cudaMemcpyAsync(ctx->d_apa_sel, sel, nsel * 8, cudaMemcpyHostToDevice, st));
    if (!init) CK(cudaMemcpyAsync(ctx->d_apa_avg, acc, (size_t)cells * 8, cudaMemcpyHostToDevice, st));
    k_apa_accumulate<<<(cells + 31) / 32, 32, 0, st>>>(ctx->d_apa_wins, ctx->d_apa_sel, nsel, cells, ctx->d_apa_avg, init ? 1 : 0);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(acc, ctx->d_apa_avg, (size_t)cells * 8, cudaMemcpyDeviceToHost, st));
    CK(stream_sync(ctx));
    return HP_OK;
}

extern "C" int hp_apa_get_windows(hp_ctx* ctx, const int64_t* sel, int64_t nsel, double* out) {
    if (!ctx || !out || nsel < 0 || (nsel && !sel)) return fail(ctx, HP_ERR_INVALID, "bad argument");
    if (!ctx->apa_npos) return fail(ctx, HP_ERR_STATE, "hp_apa_windows must come first");
    const int side = 2 * ctx->apa_w + 1, cells = side * side;
    CK(cudaSetDevice(ctx->device));
    for (int64_t k = 0; k < nsel; ++k) {
        if (sel[k] < 0 || sel[k] >= ctx->apa_npos) return fail(ctx, HP_ERR_INVALID, "window index out of range");
        CK(cudaMemcpyAsync(out + (size_t)k * cells, ctx->d_apa_wins + (size_t)sel[k] * cells, (size_t)cells * 8, cudaMemcpyDeviceToHost,
                           ctx->stream));
    }
    CK(stream_sync(ctx));
    return HP_OK;
}


// ---------------------------------------------------------------------------------------------
// Genome-wide FDR across chromosomes and GPUs (BASELINE.json's north star; NOT the reference's behaviour, which corrects
// per chromosome -- scripts/pyHICCUPS:139-198 + callers.py:263-275).  Replaces nothing in the reference; it is the one
// collective of the path: the (pair, background, lambda-chunk, observed) histograms are summed on the device over the
// contexts of this process and all-reduced over the ranks with NCCL, E.max() with a max, the valid counts with a sum.
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already holds a copy (torch's) shares it, and the
// library loads on a box without NCCL as long as nobody asks for a multi-rank communicator.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
static NcclApi* nccl_api() {
    static NcclApi* api = []() {
        NcclApi* a = new NcclApi();
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a->lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a->lib) break;
        }
        if (!a->lib) { a->err = std::string("cannot load libnccl.so.2: ") + dlerror(); return a; }
        auto sym = [&](const char* n) { void* p = dlsym(a->lib, n); if (!p) a->err = std::string("libnccl lacks ") + n; return p; };
        a->GetUniqueId = (decltype(a->GetUniqueId))sym("ncclGetUniqueId");
        a->CommInitRank = (decltype(a->CommInitRank))sym("ncclCommInitRank");
        a->CommDestroy = (decltype(a->CommDestroy))sym("ncclCommDestroy");
        a->AllReduce = (decltype(a->AllReduce))sym("ncclAllReduce");
        a->GroupStart = (decltype(a->GroupStart))sym("ncclGroupStart");
        a->GroupEnd = (decltype(a->GroupEnd))sym("ncclGroupEnd");
        a->GetErrorString = (decltype(a->GetErrorString))sym("ncclGetErrorString");
        return a;
    }();
    return api;
}
#define NCK(call)                                                                                               \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess) return fail(ctx, HP_ERR_CUDA, std::string(#call) + ": " + N->GetErrorString(r_)); \
    } while (0)

static void comm_release(hp_ctx* ctx) {
    if (ctx->comm) {
        NcclApi* N = nccl_api();
        if (N->CommDestroy) N->CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->comm_nranks = 0;
}

extern "C" int hp_comm_unique_id(void* id) {
    hp_ctx* ctx = nullptr;
    if (!id) return fail(ctx, HP_ERR_INVALID, "NULL id");
    NcclApi* N = nccl_api();
    if (!N->err.empty()) return fail(ctx, HP_ERR_CUDA, N->err);
    static_assert(sizeof(ncclUniqueId) == HP_COMM_ID_BYTES, "HP_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    NCK(N->GetUniqueId((ncclUniqueId*)id));
    return HP_OK;
}

extern "C" int hp_comm_init(hp_ctx* ctx, int32_t nranks, int32_t rank, const void* id) {
    if (!ctx) return fail(ctx, HP_ERR_INVALID, "NULL ctx");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, HP_ERR_INVALID, "need 0 <= rank < nranks");
    CK(cudaSetDevice(ctx->device));
    comm_release(ctx);
    if (nranks > 1) {
        if (!id) return fail(ctx, HP_ERR_INVALID, "a multi-rank communicator needs the unique id of rank 0 (hp_comm_unique_id)");
        NcclApi* N = nccl_api();
        if (!N->err.empty()) return fail(ctx, HP_ERR_CUDA, N->err);
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        NCK(N->CommInitRank(&ctx->comm, nranks, uid, rank));
    }
    ctx->comm_nranks = nranks; ctx->comm_rank = rank;
    return HP_OK;
}

extern "C" int hp_comm_destroy(hp_ctx* ctx) {
    if (!ctx) return fail(ctx, HP_ERR_INVALID, "NULL ctx");
    CK(cudaSetDevice(ctx->device));
    comm_release(ctx);
    return HP_OK;
}

__global__ void k_hist_accum(unsigned long long* __restrict__ acc, const unsigned int* __restrict__ h, size_t n,
                             unsigned long long* __restrict__ acc_small, const unsigned long long* __restrict__ small) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const unsigned v = h[i]; if (v) acc[i] += v; }
    if (i < 16) {                                       // [0, 16) E.max bits (max), [16, 32) valid counts (sum)
        const unsigned long long e = small[i];
        if (e > acc_small[i]) acc_small[i] = e;
        acc_small[16 + i] += small[16 + i];
    }
}
__global__ void k_hist_store(const unsigned long long* __restrict__ acc, unsigned int* __restrict__ h, size_t n,
                             const unsigned long long* __restrict__ acc_small, unsigned long long* __restrict__ small,
                             unsigned int* __restrict__ overflow) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned long long v = acc[i];
        if (v > 0xffffffffull) *overflow = 1u;
        h[i] = (unsigned)v;
    }
    if (i < 16) { small[i] = acc_small[i]; small[16 + i] = acc_small[16 + i]; }
}

extern "C" int hp_allreduce_hist(hp_ctx* ctx, hp_ctx* const* ctxs, int32_t nctx, int32_t npw_all, float* ms_out) {
    if (!ctx || (nctx > 0 && !ctxs) || nctx < 0 || npw_all < 1 || npw_all > HP_MAX_PW) return fail(ctx, HP_ERR_INVALID, "bad argument");
    if (ctx->comm_nranks < 1) return fail(ctx, HP_ERR_STATE, "hp_comm_init must come first");
    CK(cudaSetDevice(ctx->device));
    // geometry: every rank passes contexts scored with the same (pw, ww) list and chunk tables; a rank may hold none
    for (int k = 0; k < nctx; ++k) {
        hp_ctx* c = ctxs[k];
        if (!c || !c->scored) return fail(ctx, HP_ERR_STATE, "every context must have been scored (hp_hiccups_score)");
        if (c->device != ctx->device) return fail(ctx, HP_ERR_INVALID, "contexts of one call live on the communicator's GPU");
        if (c->prm.flags & HP_PF_BHFDR) return fail(ctx, HP_ERR_INVALID, "the BH-FDR caller has no lambda-chunk histograms");
        if (c->chunks.total_bins != ctx->chunks.total_bins) return fail(ctx, HP_ERR_INVALID, "contexts differ in max_chunks");
        if (c->prm.npw != npw_all) return fail(ctx, HP_ERR_INVALID, "a context was scored with a different number of (pw, ww) pairs");
    }
    cudaStream_t st = ctx->stream;
    const size_t tb = ctx->chunks.total_bins;
    // every rank reduces the same npw_all * 2 tables (a rank without a chromosome contributes zeros)
    const size_t cnt = (size_t)npw_all * 2 * tb;
    CK(ensure(&ctx->d_acc, &ctx->cap_acc, cnt + 33));
    CK(cudaMemsetAsync(ctx->d_acc, 0, (cnt + 33) * sizeof(unsigned long long), st));
    if (!ctx->ev_t0) { CK(cudaEventCreate(&ctx->ev_t0)); CK(cudaEventCreate(&ctx->ev_t1)); }
    CK(cudaEventRecord(ctx->ev[6], st));
    for (int k = 0; k < nctx; ++k) {                    // contexts of one process: plain device-side sums, in stream order
        hp_ctx* c = ctxs[k];
        const size_t n = (size_t)c->prm.npw * 2 * tb;
        k_hist_accum<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->d_acc, c->d_hist, n, ctx->d_acc + cnt, c->d_small);
        CK(cudaGetLastError());
    }
    if (ctx->comm_nranks > 1) {
        NcclApi* N = nccl_api();
        NCK(N->GroupStart());
        NCK(N->AllReduce(ctx->d_acc, ctx->d_acc, cnt, ncclUint64, ncclSum, ctx->comm, st));
        NCK(N->AllReduce(ctx->d_acc + cnt, ctx->d_acc + cnt, 16, ncclUint64, ncclMax, ctx->comm, st));
        NCK(N->AllReduce(ctx->d_acc + cnt + 16, ctx->d_acc + cnt + 16, 16, ncclUint64, ncclSum, ctx->comm, st));
        NCK(N->GroupEnd());
    }
    unsigned int* d_over = (unsigned int*)(ctx->d_acc + cnt + 32);
    for (int k = 0; k < nctx; ++k) {
        hp_ctx* c = ctxs[k];
        const size_t n = (size_t)c->prm.npw * 2 * tb;
        k_hist_store<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->d_acc, c->d_hist, n, ctx->d_acc + cnt, c->d_small, d_over);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[7], st));
    unsigned long long small[33];
    CK(cudaMemcpyAsync(ctx->h_res + 6400, ctx->d_acc + cnt, sizeof(small), cudaMemcpyDeviceToHost, st));
    CK(stream_sync(ctx));
    memcpy(small, ctx->h_res + 6400, sizeof(small));
    if ((unsigned int)small[32]) return fail(ctx, HP_ERR_CAPACITY, "a merged histogram bin exceeds 2^32 - 1");
    for (int k = 0; k < nctx; ++k) {
        hp_ctx* c = ctxs[k];
        for (int i = 0; i < c->prm.npw; ++i)
            for (int fl = 0; fl < 2; ++fl) {
                hp_lf_stat& L = c->sum.lf[i][fl];
                L.n_valid = (int64_t)small[16 + i * 2 + fl];
                double em;
                memcpy(&em, &small[i * 2 + fl], 8);
                L.e_max = em;
                L.numbin = (L.n_valid > 0) ? (int)ceil(log(em) / log(2.0) * 3 + 1) : 0;
            }
        c->fdr_done = false;
    }
    if (ms_out) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); *ms_out = ms; }
    return HP_OK;
}

// upload scratch (device landing zone, pinned staging, prep tables) is only needed during an upload: a caller that
// keeps many scored contexts alive (genome-wide FDR: one per chromosome until the merged BH) gives it back
extern "C" int hp_ctx_trim(hp_ctx* ctx) {
    if (!ctx) return fail(ctx, HP_ERR_INVALID, "NULL ctx");
    CK(cudaSetDevice(ctx->device));
    CK(stream_sync(ctx));
    if (ctx->d_tmp) { cudaFree(ctx->d_tmp); ctx->d_tmp = nullptr; ctx->cap_plane = 0; }
    // cap_plane guards d_raw / d_bal / d_lvl as well: the next upload reallocates them (the band on the device is kept
    // valid until then)
    if (ctx->h_stage) { cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; ctx->cap_stage = 0; }
    if (ctx->d_prep) { cudaFree(ctx->d_prep); ctx->d_prep = nullptr; ctx->cap_prep = 0; }
    if (ctx->d_dump) { cudaFree(ctx->d_dump); ctx->d_dump = nullptr; ctx->cap_dump = 0; }
    return HP_OK;
}
