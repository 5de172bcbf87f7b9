// The worker's inner loop (worker_bwt.erl:73-108) as a resident batch pipeline, and the C ABI on top of it.
//
//   upload (H2D: ASCII + offsets, or 2-bit packed words)  ->  [pack]  ->  sga_kernel  ->  sga_slow_kernel  ->  windows
//                 ->  sw_classify / sw_wfa / sw_fast  ->  emit (scan + gather into the caller-facing layout)  ->  download (D2H)
//
// A run is ENQUEUED as a whole: every kernel takes its counts from device memory and every buffer is sized by capacity,
// so the host synchronises once, at the end, and then only looks at a handful of counters.  What the data rarely needs
// (wide windows, rows that do not fit the difference coding, strings with 'N', capacity overflows) is a second,
// synchronous round behind that.  One bwt_batch owns its stream, device buffers and pinned staging, so several batches
// can be in flight on one (immutable) index from different host threads, and their kernels overlap on the device.
#include <algorithm>
#include <cstddef>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>

#include "common.cuh"
#include "reads.cuh"
#include "seed.cuh"
#include "sw.cuh"

namespace {

template <typename T>
int dev_grow(T *&p, size_t &cap, size_t need, bool keep_slack = true) {
    if (p && need <= cap) return BWT_OK;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = keep_slack ? need + need / 4 + 64 : need;
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        cap = 0;
        bwt_set_error("device allocation of %zu bytes failed", want * sizeof(T));
        return BWT_ERR_NOMEM;
    }
    cap = want;
    return BWT_OK;
}

// pipelined aligner: results of a chunk move to their place in the caller's arrays
__global__ void rebase_result_kernel(int64_t *__restrict__ hit_off64, int64_t n, bwt_hit *__restrict__ hits, int64_t n_hits,
                                     int64_t hit_base, int64_t cig_base) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) hit_off64[k] += hit_base;
    if (k < n_hits) hits[k].cigar_off += cig_base;
}
__global__ void rebase_result16_kernel(uint32_t *__restrict__ hit_off, int64_t n, bwt_hit16 *__restrict__ hits, int64_t n_hits,
                                       uint32_t hit_base, uint32_t cig_base) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) hit_off[k] += hit_base;
    if (k < n_hits) hits[k].cigar_off += cig_base;
}

// windows of worker_bwt.erl:84-90: Ref_len = Lq + D, Start = S - Ref_len + 200 into the padded reference
__global__ void windows_kernel(const Cluster *__restrict__ cl, const uint32_t *__restrict__ n_ptr, const uint32_t *__restrict__ len,
                               uint64_t ref_len_padded, SwDesc *__restrict__ desc, SwOut *__restrict__ out) {
    const uint32_t n = *n_ptr;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        Cluster c = cl[k];
        const int32_t lq = (int32_t)len[c.read];
        const int64_t ref_len = (int64_t)lq + c.d;
        const int64_t start = (int64_t)c.s - ref_len + BWT_REF_EXT;
        SwDesc d;
        d.r_off = start; d.read = c.read; d.lq = lq; d.lr = (int32_t)ref_len; d.pad = 0;
        desc[k] = d;
        SwOut o;
        o.score = 0; o.cigar_len = 0; o.cigar_off = 0;
        // <<_:Start_pos/bytes, Ref_seq:Ref_len/bytes, _/binary>> = Ref  -> badmatch outside the binary
        o.status = (start < 0 || (uint64_t)(start + ref_len) > ref_len_padded) ? BWT_HIT_WINDOW_OOR : -1;
        out[k] = o;
    }
}

// per read: number of hits and CIGAR bytes (inputs of the two scans).  compact: "<Lq>M" CIGARs are not stored.
__global__ void emit_count_kernel(const uint32_t *__restrict__ ncl, const uint32_t *__restrict__ cl_base,
                                  const SwOut *__restrict__ out, int64_t n, int compact, const uint32_t *__restrict__ abort,
                                  uint32_t *__restrict__ hit_cnt, uint32_t *__restrict__ cig_cnt) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n || *abort) { hit_cnt[r] = 0; cig_cnt[r] = 0; return; }   // slot n turns into the total under the exclusive scan
    uint32_t c = ncl[r], b = cl_base[r], bytes = 0;
    for (uint32_t k = 0; k < c; k++) {
        const SwOut o = out[b + k];
        if (!(compact && (o.status & SW_FULL_M))) bytes += (uint32_t)o.cigar_len;
    }
    hit_cnt[r] = c;
    cig_cnt[r] = bytes;
}

// per read: hits in ASCENDING position (the reference prepends while folding sga's descending list,
// worker_bwt.erl:84-97), CIGAR bytes packed in the same order
__global__ void emit_kernel(const uint32_t *__restrict__ ncl, const uint32_t *__restrict__ cl_base,
                            const Cluster *__restrict__ cl, const SwDesc *__restrict__ desc,
                            const SwOut *__restrict__ out, const char *__restrict__ cig_tmp, int64_t n,
                            const uint32_t *__restrict__ hit_off, const uint32_t *__restrict__ cig_off, int64_t shift,
                            int64_t *__restrict__ hit_off64, bwt_hit *__restrict__ hits, char *__restrict__ cigars,
                            uint32_t hits_cap, uint32_t cig_cap, uint32_t tmp_cap, const uint32_t *__restrict__ abort) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n || *abort) { hit_off64[r] = hit_off[r]; return; }
    uint32_t c = ncl[r], b = cl_base[r], ho = hit_off[r], co = cig_off[r];
    hit_off64[r] = ho;
    for (uint32_t k = 0; k < c; k++) {
        uint32_t src = b + (c - 1 - k);
        Cluster cc = cl[src];
        SwOut o = out[src];
        bwt_hit h;
        h.s = cc.s;
        h.d = cc.d;
        h.pos = (int64_t)cc.s - (int64_t)desc[src].lr + shift;     // S - Ref_len + Shift (worker_bwt.erl:94)
        h.score = o.score;
        h.status = o.status & 0xFF;
        h.cigar_len = o.cigar_len;
        h.cigar_off = co;
        if (ho + k < hits_cap) hits[ho + k] = h;
        if ((uint64_t)co + (uint32_t)o.cigar_len <= cig_cap && (uint64_t)o.cigar_off + (uint32_t)o.cigar_len <= tmp_cap)
            for (int t = 0; t < o.cigar_len; t++) cigars[co + t] = cig_tmp[o.cigar_off + t];
        co += (uint32_t)o.cigar_len;
    }
}

// the same in the compact wire format (bwt_hit16, uint8 status, uint32 offsets)
__global__ void emit16_kernel(const uint32_t *__restrict__ ncl, const uint32_t *__restrict__ cl_base,
                              const Cluster *__restrict__ cl, const SwOut *__restrict__ out, const char *__restrict__ cig_tmp,
                              const int32_t *__restrict__ read_status, int64_t n, const uint32_t *__restrict__ hit_off,
                              const uint32_t *__restrict__ cig_off, uint8_t *__restrict__ status8, bwt_hit16 *__restrict__ hits,
                              char *__restrict__ cigars, uint32_t hits_cap, uint32_t cig_cap, uint32_t tmp_cap,
                              const uint32_t *__restrict__ abort, uint32_t *__restrict__ flags) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || *abort) return;
    status8[r] = (uint8_t)read_status[r];
    uint32_t c = ncl[r], b = cl_base[r], ho = hit_off[r], co = cig_off[r];
    for (uint32_t k = 0; k < c; k++) {
        uint32_t src = b + (c - 1 - k);
        Cluster cc = cl[src];
        SwOut o = out[src];
        const bool full_m = (o.status & SW_FULL_M) != 0;
        if (o.cigar_len > 65535 || (uint32_t)cc.d >= (1u << 28) || o.score > 32767 || o.score < -32768) atomicExch(flags, 1u);
        bwt_hit16 h;
        h.s = cc.s;
        h.d_flags = ((uint32_t)cc.d & 0x0FFFFFFFu) | ((uint32_t)(o.status & 3) << 28) | (full_m ? BWT_HIT16_FULL_M : 0u);
        h.score = (int16_t)o.score;
        h.cigar_len = full_m ? (uint16_t)0 : (uint16_t)o.cigar_len;
        h.cigar_off = full_m ? 0u : co;
        if (ho + k < hits_cap) hits[ho + k] = h;
        if (!full_m) {
            if ((uint64_t)co + (uint32_t)o.cigar_len <= cig_cap && (uint64_t)o.cigar_off + (uint32_t)o.cigar_len <= tmp_cap)
                for (int t = 0; t < o.cigar_len; t++) cigars[co + t] = cig_tmp[o.cigar_off + t];
            co += (uint32_t)o.cigar_len;
        }
    }
}

// sga-only result in the reference's order (descending S)
__global__ void emit_sga_kernel(const uint32_t *__restrict__ ncl, const uint32_t *__restrict__ cl_base,
                                const Cluster *__restrict__ cl, int64_t n, const uint32_t *__restrict__ off,
                                int64_t *__restrict__ off64, int64_t *__restrict__ S, int32_t *__restrict__ D) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    off64[r] = off[r];
    if (r == n) return;
    uint32_t c = ncl[r], b = cl_base[r], o = off[r];
    for (uint32_t k = 0; k < c; k++) { S[o + k] = cl[b + k].s; D[o + k] = cl[b + k].d; }
}

__global__ void desc_from_offsets_kernel(const int64_t *__restrict__ qo, const int64_t *__restrict__ ro, uint32_t n,
                                         SwDesc *__restrict__ desc, SwOut *__restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    SwDesc d;
    d.r_off = ro[k]; d.read = k; d.lq = (int32_t)(qo[k + 1] - qo[k]); d.lr = (int32_t)(ro[k + 1] - ro[k]); d.pad = 0;
    desc[k] = d;
    SwOut o; o.score = 0; o.status = -1; o.cigar_len = 0; o.cigar_off = 0;
    out[k] = o;
}

__global__ void sw_pairs_emit_count_kernel(const SwOut *__restrict__ out, uint32_t n, uint32_t *__restrict__ cnt) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) cnt[k] = (uint32_t)out[k].cigar_len;
}

__global__ void sw_pairs_emit_kernel(const SwOut *__restrict__ out, const char *__restrict__ cig_tmp, uint32_t n,
                                     const uint32_t *__restrict__ off, int32_t *__restrict__ status,
                                     int32_t *__restrict__ score, int64_t *__restrict__ cig_off,
                                     int32_t *__restrict__ cig_len, char *__restrict__ cigars) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    SwOut o = out[k];
    status[k] = o.status & 0xFF; score[k] = o.score; cig_off[k] = off[k]; cig_len[k] = o.cigar_len;
    for (int t = 0; t < o.cigar_len; t++) cigars[off[k] + t] = cig_tmp[o.cigar_off + t];
}

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

// what the last lines of a run copy back, in one pinned block
struct RunTotals {
    uint32_t hits, cig_bytes;       // exclusive-scan totals of the emit stage
    uint32_t info[2];               // [1]: the ASCII offsets decreased
    uint32_t emit_flags;            // compact emit: a field did not fit bwt_hit16
    uint32_t pad;
};

struct bwt_batch {
    bwt_index *ix = nullptr;
    cudaStream_t stream = nullptr;
    int64_t max_reads = 0, max_bytes = 0;
    // reads
    uint8_t *d_ascii = nullptr;  size_t ascii_cap = 0;
    int64_t *d_aoff = nullptr;
    uint32_t *d_pk = nullptr;    size_t pk_cap = 0;        // words
    uint32_t *d_woff = nullptr, *d_len = nullptr, *d_exc = nullptr, *d_exc_idx = nullptr;
    size_t exc_cap = 0;
    uint16_t *d_len16 = nullptr;
    DevReads R{};
    bool need_pack = false;               // ASCII batch: the run packs it first
    uint8_t *h_stage = nullptr;           // pinned staging for pageable callers
    size_t h_stage_bytes = 0;
    int64_t n = 0, n_bytes = 0;
    int max_read_len = 0;
    // per read
    int32_t *d_read_status = nullptr;
    uint32_t *d_ncl = nullptr, *d_cl_base = nullptr, *d_hit_off = nullptr, *d_cig_off = nullptr, *d_scan_tmp = nullptr;
    int64_t *d_hit_off64 = nullptr;
    uint8_t *d_status8 = nullptr;
    // clusters
    Cluster *d_clusters = nullptr;        // dense list (what windows / emit read)
    Cluster *d_clusters_sparse = nullptr; // the sub-list regions sga_kernel appends to (launch_sga_finish packs them)
    size_t cl_cap = 0, cl_sparse_cap = 0;
    SlowRead *d_slow = nullptr;
    uint32_t slow_cap = 0;
    uint32_t *d_slow_scratch = nullptr;
    size_t slow_scratch_cap = 0;
    SeedCounters *d_ctr = nullptr, *h_ctr = nullptr;
    RunTotals *d_totals = nullptr, *h_totals = nullptr;
    SwStage sw;
    // outputs (sized by capacity: cl_cap hits, sw.cig_cap CIGAR bytes)
    bwt_hit *d_hits = nullptr;     size_t hits_cap = 0;
    bwt_hit16 *d_hits16 = nullptr; size_t hits16_cap = 0;
    char *d_cigars = nullptr;      size_t cigars_cap = 0;
    int64_t n_hits = 0, cigar_bytes = 0;
    bool has_result = false, result_compact = false;
    cudaEvent_t ev[10] = {};
    bwt_stats stats{};
    cudaStream_t dl_stream = nullptr;     // the aligner's D2H copies: the next chunk's upload does not wait for them
    cudaEvent_t ev_run_done = nullptr, ev_dl0 = nullptr, ev_dl1 = nullptr;
    cudaEvent_t ev_block = nullptr, ev_block_dl = nullptr;      // cudaEventBlockingSync: see batch_wait
    bool dl_pending = false;
};

// Wait for everything enqueued on `stream`.  Default: cudaStreamSynchronize (the driver spins: lowest latency; blocking
// waits cost 8 % of the e2e rate on one GPU, 484 -> 443 M reads/s).  BWT_BLOCKING_SYNC=1: large batches wait on an event
// created with cudaEventBlockingSync, the calling thread sleeps.  A pipelined aligner keeps `depth` such threads per GPU;
// on a node where ranks x (depth + 1) exceeds the host cores the spinning threads starve each other and the threads that
// should be enqueueing the next chunk, and the host application (bench.py does) turns the blocking wait on.
static cudaError_t batch_wait(bwt_batch *b, cudaStream_t stream, cudaEvent_t ev_block, bool large) {
    static int mode = -2;
    if (mode == -2) { const char *e = getenv("BWT_BLOCKING_SYNC"); mode = e && *e ? atoi(e) : -1; }
    const bool block = mode > 0 && large;
    (void)b;
    if (!block || !ev_block) return cudaStreamSynchronize(stream);
    cudaError_t e = cudaEventRecord(ev_block, stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev_block);
}

static int batch_alloc(bwt_batch *b) {
    int64_t nr = b->max_reads;
    CU_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&b->dl_stream, cudaStreamNonBlocking));
    for (auto &e : b->ev) CU_TRY(cudaEventCreate(&e));
    CU_TRY(cudaEventCreate(&b->ev_run_done));
    CU_TRY(cudaEventCreate(&b->ev_dl0));
    CU_TRY(cudaEventCreate(&b->ev_dl1));
    CU_TRY(cudaEventCreateWithFlags(&b->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&b->ev_block_dl, cudaEventBlockingSync | cudaEventDisableTiming));
    b->pk_cap = (size_t)b->max_bytes / 16 + (size_t)nr + 2;
    CU_TRY(cudaMalloc(&b->d_pk, b->pk_cap * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_aoff, (size_t)(nr + 1) * sizeof(int64_t)));
    CU_TRY(cudaMalloc(&b->d_woff, (size_t)(nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_len, (size_t)(nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_len16, (size_t)(nr + 1) * sizeof(uint16_t)));
    CU_TRY(cudaMalloc(&b->d_exc, (size_t)(nr / 32 + 2) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_read_status, (size_t)nr * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&b->d_status8, (size_t)nr));
    CU_TRY(cudaMalloc(&b->d_ncl, (size_t)nr * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_cl_base, (size_t)nr * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_hit_off, (size_t)(nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_cig_off, (size_t)(nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_hit_off64, (size_t)(nr + 1) * sizeof(int64_t)));
    CU_TRY(cudaMalloc(&b->d_scan_tmp, scan_tmp_words((size_t)nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_ctr, sizeof(SeedCounters)));
    CU_TRY(cudaMallocHost(&b->h_ctr, sizeof(SeedCounters)));
    CU_TRY(cudaMalloc(&b->d_totals, sizeof(RunTotals)));
    CU_TRY(cudaMallocHost(&b->h_totals, sizeof(RunTotals)));
    b->slow_cap = (uint32_t)std::min<int64_t>(nr, 1 << 16);
    CU_TRY(cudaMalloc(&b->d_slow, (size_t)b->slow_cap * sizeof(SlowRead)));
    BWT_TRY(dev_grow(b->d_slow_scratch, b->slow_scratch_cap, (size_t)1 << 20, false));
    b->sw.device = b->ix->device;
    size_t want_cl = (size_t)std::max<int64_t>(4 * nr, 4096);
    BWT_TRY(dev_grow(b->d_clusters, b->cl_cap, want_cl, false));
    BWT_TRY(dev_grow(b->d_clusters_sparse, b->cl_sparse_cap, b->cl_cap, false));
    BWT_TRY(sw_stage_reserve(b->sw, (uint32_t)b->cl_cap));
    return BWT_OK;
}

extern "C" int bwt_batch_create(bwt_index *h, int64_t max_reads, int64_t max_read_bytes, bwt_batch **out) {
    if (!h || !out || max_reads <= 0 || max_read_bytes <= 0 || max_reads >= (1ll << 31)) {
        bwt_set_error("bwt_batch_create: bad argument");
        return BWT_ERR_ARG;
    }
    BWT_TRY(index_select_device(h->device));
    bwt_batch *b = new bwt_batch();
    b->ix = h;
    b->max_reads = max_reads;
    b->max_bytes = max_read_bytes;
    int st = batch_alloc(b);
    if (st != BWT_OK) { bwt_batch_destroy(b); return st; }
    *out = b;
    return BWT_OK;
}

extern "C" int bwt_batch_destroy(bwt_batch *b) {
    if (!b) return BWT_OK;
    cudaSetDevice(b->ix->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->dl_stream) cudaStreamSynchronize(b->dl_stream);
    cudaFree(b->d_ascii); cudaFree(b->d_aoff); cudaFree(b->d_pk); cudaFree(b->d_woff); cudaFree(b->d_len); cudaFree(b->d_len16);
    cudaFree(b->d_exc); cudaFree(b->d_exc_idx); cudaFree(b->d_read_status); cudaFree(b->d_status8); cudaFree(b->d_ncl); cudaFree(b->d_cl_base);
    cudaFree(b->d_hit_off); cudaFree(b->d_cig_off); cudaFree(b->d_hit_off64); cudaFree(b->d_scan_tmp);
    cudaFree(b->d_clusters); cudaFree(b->d_clusters_sparse); cudaFree(b->d_slow); cudaFree(b->d_slow_scratch);
    cudaFree(b->d_ctr); cudaFree(b->d_totals); cudaFree(b->d_hits); cudaFree(b->d_hits16); cudaFree(b->d_cigars);
    if (b->h_ctr) cudaFreeHost(b->h_ctr);
    if (b->h_totals) cudaFreeHost(b->h_totals);
    if (b->h_stage) cudaFreeHost(b->h_stage);
    sw_stage_free(b->sw);
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    if (b->ev_run_done) cudaEventDestroy(b->ev_run_done);
    if (b->ev_dl0) cudaEventDestroy(b->ev_dl0);
    if (b->ev_block) cudaEventDestroy(b->ev_block);
    if (b->ev_block_dl) cudaEventDestroy(b->ev_block_dl);
    if (b->ev_dl1) cudaEventDestroy(b->ev_dl1);
    if (b->dl_stream) cudaStreamDestroy(b->dl_stream);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return BWT_OK;
}

static int stage_reserve(bwt_batch *b, size_t need) {
    if (need <= b->h_stage_bytes) return BWT_OK;
    if (b->h_stage) cudaFreeHost(b->h_stage);
    b->h_stage = nullptr;
    b->h_stage_bytes = need + need / 4 + 4096;
    CU_TRY(cudaMallocHost(&b->h_stage, b->h_stage_bytes));
    return BWT_OK;
}

static void view_reset(bwt_batch *b) {
    DevReads &R = b->R;
    R.pk = b->d_pk; R.woff = b->d_woff; R.len = b->d_len; R.exc = b->d_exc;
    R.ascii = b->d_ascii; R.aoff = b->d_aoff; R.exc_idx = b->d_exc_idx;
    R.n_exc = 0; R.ascii_full = 0; R.aoff_base = 0; R.exc_base = 0;
}

// ASCII batch: the bytes and the offsets go up as they are; lengths, word offsets and the 2-bit words are made on the
// device (the packing itself at the start of the run: it is part of the work on an ASCII batch).  No synchronisation
// unless the caller's memory is pageable (then it is staged through pinned memory, which must not be reused too early).
extern "C" int bwt_batch_upload(bwt_batch *b, const uint8_t *reads, const int64_t *offsets, int64_t n) {
    if (!b || !offsets || n < 0 || (n > 0 && !reads)) { bwt_set_error("bwt_batch_upload: bad argument"); return BWT_ERR_ARG; }
    if (n > b->max_reads) { bwt_set_error("batch of %lld reads exceeds max_reads %lld", (long long)n, (long long)b->max_reads); return BWT_ERR_ARG; }
    const int64_t base = offsets[0], bytes = offsets[n] - base;
    if (bytes < 0 || bytes > b->max_bytes) { bwt_set_error("batch of %lld read bytes exceeds max_read_bytes %lld", (long long)bytes, (long long)b->max_bytes); return BWT_ERR_ARG; }
    // the longest read and the monotonic check on the host: the offsets are here anyway
    int64_t maxlen = 0, words = 0;
    for (int64_t k = 0; k < n; k++) {
        const int64_t l = offsets[k + 1] - offsets[k];
        if (l < 0) { bwt_set_error("offsets must be non-decreasing"); return BWT_ERR_ARG; }
        maxlen = std::max(maxlen, l);
        if (l <= BWT_PACK_MAX) words += (l + 15) >> 4;
    }
    CU_TRY(cudaSetDevice(b->ix->device));
    if ((size_t)words + 2 > b->pk_cap) BWT_TRY(dev_grow(b->d_pk, b->pk_cap, (size_t)words + 2));
    if ((size_t)bytes + 16 > b->ascii_cap) BWT_TRY(dev_grow(b->d_ascii, b->ascii_cap, (size_t)bytes + 16, false));
    const bool pin_r = bytes == 0 || is_pinned(reads + base), pin_o = is_pinned(offsets);
    BWT_TRY(stage_reserve(b, (pin_r ? 0 : (size_t)bytes) + (pin_o ? 0 : (size_t)(n + 1) * sizeof(int64_t))));
    const int64_t *src_o = offsets;
    if (!pin_o) { memcpy(b->h_stage, offsets, (size_t)(n + 1) * sizeof(int64_t)); src_o = reinterpret_cast<const int64_t *>(b->h_stage); }
    CU_TRY(cudaEventRecord(b->ev[4], b->stream));
    CU_TRY(cudaMemcpyAsync(b->d_aoff, src_o, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, b->stream));
    if (bytes > 0) {
        const uint8_t *src = reads + base;
        if (!pin_r) {
            uint8_t *sr = b->h_stage + (pin_o ? 0 : (size_t)(n + 1) * sizeof(int64_t));
            memcpy(sr, src, (size_t)bytes);
            src = sr;
        }
        CU_TRY(cudaMemcpyAsync(b->d_ascii, src, (size_t)bytes, cudaMemcpyHostToDevice, b->stream));
    }
    CU_TRY(cudaEventRecord(b->ev[5], b->stream));
    view_reset(b);
    b->R.ascii_full = 1; b->R.aoff_base = base;
    b->need_pack = true;
    b->n = n; b->n_bytes = bytes; b->max_read_len = (int)std::min<int64_t>(maxlen, 1 << 30);
    b->has_result = false;
    if (!pin_r || !pin_o) CU_TRY(cudaStreamSynchronize(b->stream));     // the staging buffer is free again
    return BWT_OK;
}

// Packed batch (bwt_pack_reads / a FASTQ ingest that packs): words, lengths and the flagged reads' ASCII go up; a
// quarter of the bytes of an ASCII batch.  lo/hi select reads [lo, hi) of p (the pipelined aligner's chunks);
// word_base = words of p before read lo.
static int batch_upload_packed_range(bwt_batch *b, const bwt_packed_reads *p, int64_t lo, int64_t hi, int64_t word_base,
                                     int64_t words, int maxlen) {
    const int64_t n = hi - lo;
    if (n > b->max_reads) { bwt_set_error("batch of %lld reads exceeds max_reads %lld", (long long)n, (long long)b->max_reads); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(b->ix->device));
    if ((size_t)words + 2 > b->pk_cap) BWT_TRY(dev_grow(b->d_pk, b->pk_cap, (size_t)words + 2));
    // the flagged reads of the range
    int64_t e0 = 0, e1 = 0;
    if (p->n_exc > 0) {
        e0 = std::lower_bound(p->exc_idx, p->exc_idx + p->n_exc, (uint32_t)lo) - p->exc_idx;
        e1 = std::lower_bound(p->exc_idx, p->exc_idx + p->n_exc, (uint32_t)hi) - p->exc_idx;
    }
    const int64_t ne = e1 - e0, eb = ne ? p->exc_off[e1] - p->exc_off[e0] : 0;
    if ((size_t)ne + 1 > b->exc_cap) BWT_TRY(dev_grow(b->d_exc_idx, b->exc_cap, (size_t)ne + 1));
    if ((size_t)eb + 16 > b->ascii_cap) BWT_TRY(dev_grow(b->d_ascii, b->ascii_cap, (size_t)eb + 16));
    if (ne > b->max_reads) { bwt_set_error("more flagged reads than reads"); return BWT_ERR_ARG; }
    const bool pinned = is_pinned(p->words) && (p->fixed_len > 0 || is_pinned(p->lens)) && (ne == 0 || (is_pinned(p->exc_idx) && is_pinned(p->exc_off) && is_pinned(p->exc_ascii)));
    const uint32_t *src_w = p->words + word_base;
    const uint16_t *src_l = p->fixed_len > 0 ? nullptr : p->lens + lo;
    const uint32_t *src_ei = ne ? p->exc_idx + e0 : nullptr;
    const int64_t *src_eo = ne ? p->exc_off + e0 : nullptr;
    const uint8_t *src_ea = ne ? p->exc_ascii + p->exc_off[e0] : nullptr;
    if (!pinned) {
        const size_t bw = (size_t)(words + 1) * 4, bl = src_l ? (size_t)n * 2 : 0, bi = (size_t)ne * 4, bo = (size_t)(ne ? ne + 1 : 0) * 8;
        BWT_TRY(stage_reserve(b, bw + bl + bi + bo + (size_t)eb + 64));
        uint8_t *s = b->h_stage;
        memcpy(s, src_w, bw); src_w = (const uint32_t *)s; s += (bw + 7) & ~(size_t)7;
        if (ne) { memcpy(s, src_eo, bo); src_eo = (const int64_t *)s; s += bo; }
        if (ne) { memcpy(s, src_ei, bi); src_ei = (const uint32_t *)s; s += (bi + 7) & ~(size_t)7; }
        if (src_l) { memcpy(s, src_l, bl); src_l = (const uint16_t *)s; s += (bl + 7) & ~(size_t)7; }
        if (ne) { memcpy(s, src_ea, (size_t)eb); src_ea = s; }
    }
    CU_TRY(cudaEventRecord(b->ev[4], b->stream));
    CU_TRY(cudaMemcpyAsync(b->d_pk, src_w, (size_t)(words + 1) * 4, cudaMemcpyHostToDevice, b->stream));   // + the pad word
    if (src_l) CU_TRY(cudaMemcpyAsync(b->d_len16, src_l, (size_t)n * 2, cudaMemcpyHostToDevice, b->stream));
    if (ne) {
        CU_TRY(cudaMemcpyAsync(b->d_exc_idx, src_ei, (size_t)ne * 4, cudaMemcpyHostToDevice, b->stream));
        CU_TRY(cudaMemcpyAsync(b->d_aoff, src_eo, (size_t)(ne + 1) * 8, cudaMemcpyHostToDevice, b->stream));
        CU_TRY(cudaMemcpyAsync(b->d_ascii, src_ea, (size_t)eb, cudaMemcpyHostToDevice, b->stream));
    }
    CU_TRY(cudaEventRecord(b->ev[5], b->stream));
    uint32_t launches = 0;
    if (p->fixed_len > 0) BWT_TRY(reads_geom_fixed(n, (uint32_t)p->fixed_len, b->d_len, b->d_woff, b->stream, &launches));
    else BWT_TRY(reads_geom_from_len16(b->d_len16, n, b->d_len, b->d_woff, b->d_scan_tmp, b->stream, &launches));
    view_reset(b);
    b->R.n_exc = (uint32_t)ne; b->R.exc_base = (uint32_t)lo; b->R.aoff_base = ne ? p->exc_off[e0] : 0;
    BWT_TRY(reads_exc_scatter(b->d_exc_idx, (uint32_t)ne, (uint32_t)lo, n, b->d_exc, b->stream, &launches));
    b->need_pack = false;
    b->n = n; b->n_bytes = words * 4; b->max_read_len = maxlen;
    b->has_result = false;
    if (!pinned) CU_TRY(cudaStreamSynchronize(b->stream));
    return BWT_OK;
}

static int packed_check(const bwt_packed_reads *p) {
    if (!p || p->n < 0 || p->n >= (1ll << 31) || (p->n > 0 && !p->words) || (p->fixed_len <= 0 && p->n > 0 && !p->lens) || p->n_exc < 0 ||
        (p->n_exc > 0 && (!p->exc_idx || !p->exc_off || !p->exc_ascii)) || p->fixed_len > 65535) {
        bwt_set_error("bwt_packed_reads: bad argument");
        return BWT_ERR_ARG;
    }
    return BWT_OK;
}

// words before read `at`, and the longest read of [lo, hi), of a packed batch: from its index where there is one
static int64_t packed_words_before(const bwt_packed_reads *p, int64_t at) {
    if (p->fixed_len > 0) return at * (p->fixed_len > BWT_PACK_MAX ? 0 : (p->fixed_len + 15) >> 4);
    int64_t w = 0, from = 0;
    if (p->word_index) { from = at / BWT_PACK_INDEX_STRIDE * BWT_PACK_INDEX_STRIDE; w = p->word_index[at / BWT_PACK_INDEX_STRIDE]; }
    for (int64_t r = from; r < at; r++) {
        const int l = p->lens[r];
        if (l <= BWT_PACK_MAX) w += (l + 15) >> 4;
    }
    return w;
}
static int packed_maxlen(const bwt_packed_reads *p, int64_t lo, int64_t hi) {
    if (p->fixed_len > 0) return p->fixed_len;
    int m = 0;
    int64_t r = lo;
    while (r < hi) {
        if (p->maxlen_index && r % BWT_PACK_INDEX_STRIDE == 0 && r + BWT_PACK_INDEX_STRIDE <= hi) {
            m = std::max<int>(m, p->maxlen_index[r / BWT_PACK_INDEX_STRIDE]);
            r += BWT_PACK_INDEX_STRIDE;
        } else { m = std::max<int>(m, p->lens[r]); r++; }
    }
    return m;
}
static void packed_geom(const bwt_packed_reads *p, int64_t lo, int64_t hi, int64_t *words, int *maxlen) {
    *words = packed_words_before(p, hi) - packed_words_before(p, lo);
    *maxlen = packed_maxlen(p, lo, hi);
}

extern "C" int bwt_batch_upload_packed(bwt_batch *b, const bwt_packed_reads *p) {
    if (!b) { bwt_set_error("bwt_batch_upload_packed: null batch"); return BWT_ERR_ARG; }
    BWT_TRY(packed_check(p));
    int64_t words;
    int maxlen;
    packed_geom(p, 0, p->n, &words, &maxlen);
    return batch_upload_packed_range(b, p, 0, p->n, 0, words, maxlen);
}

// ---- one run ------------------------------------------------------------------------------------------------------

// everything up to the dense cluster list, enqueued
static int enqueue_seed_stage(bwt_batch *b, uint32_t cl_cap) {
    const DevIndex &ix = b->ix->dv;
    CU_TRY(cudaMemsetAsync(b->d_ctr, 0, sizeof(SeedCounters), b->stream));
    CU_TRY(cudaEventRecord(b->ev[6], b->stream));
    BWT_TRY(launch_sga(ix, b->R, b->n, b->max_read_len, b->d_read_status, b->d_ncl, b->d_cl_base,
                       b->d_clusters_sparse, cl_cap, b->d_slow, b->slow_cap, b->d_ctr, b->stream));
    CU_TRY(cudaEventRecord(b->ev[7], b->stream));
    // reads whose seed ends did not fit on chip (a few blocks that leave at once when there are none)
    BWT_TRY(launch_sga_slow(ix, b->R, b->d_slow, b->slow_cap, b->d_slow_scratch, b->slow_scratch_cap, b->d_read_status,
                            b->d_ncl, b->d_cl_base, b->d_clusters_sparse, cl_cap, b->d_ctr, b->stream));
    b->stats.launches += 2;
    // sub-lists -> dense cluster list, dense cl_base, n_clusters
    return launch_sga_finish(b->d_clusters_sparse, b->d_clusters, cl_cap, b->d_cl_base, b->d_ncl, b->n, b->slow_cap, b->d_ctr, b->stream,
                             &b->stats.launches);
}

static int enqueue_emit_stage(bwt_batch *b, bool compact) {
    const int64_t n = b->n;
    const unsigned blocks = (unsigned)((n + 1 + 255) / 256);
    const size_t hcap = b->cl_cap, ccap = (size_t)b->sw.cig_cap + 16;
    if (compact) BWT_TRY(dev_grow(b->d_hits16, b->hits16_cap, hcap, false));
    else BWT_TRY(dev_grow(b->d_hits, b->hits_cap, hcap, false));
    BWT_TRY(dev_grow(b->d_cigars, b->cigars_cap, ccap, false));
    emit_count_kernel<<<blocks, 256, 0, b->stream>>>(b->d_ncl, b->d_cl_base, b->sw.d_out, n, compact ? 1 : 0, &b->d_ctr->abort, b->d_hit_off, b->d_cig_off);
    b->stats.launches += 1;
    K_CHECK(b->stream, "emit_count_kernel");
    // element n of an exclusive scan over n+1 items is the total (item n is written as 0 above)
    BWT_TRY(exclusive_scan_u32(b->d_hit_off, (size_t)n + 1, b->d_scan_tmp, &b->d_totals->hits, b->stream, &b->stats.launches));
    BWT_TRY(exclusive_scan_u32(b->d_cig_off, (size_t)n + 1, b->d_scan_tmp, &b->d_totals->cig_bytes, b->stream, &b->stats.launches));
    if (compact)
        emit16_kernel<<<blocks, 256, 0, b->stream>>>(b->d_ncl, b->d_cl_base, b->d_clusters, b->sw.d_out, b->sw.d_cig_tmp, b->d_read_status, n,
                                                     b->d_hit_off, b->d_cig_off, b->d_status8, b->d_hits16, b->d_cigars,
                                                     (uint32_t)std::min<size_t>(b->hits16_cap, 0xFFFFFFFFu), (uint32_t)std::min<size_t>(b->cigars_cap, 0xFFFFFFFFu),
                                                     b->sw.cig_cap, &b->d_ctr->abort, &b->d_totals->emit_flags);
    else
        emit_kernel<<<blocks, 256, 0, b->stream>>>(b->d_ncl, b->d_cl_base, b->d_clusters, b->sw.d_desc, b->sw.d_out,
                                                   b->sw.d_cig_tmp, n, b->d_hit_off, b->d_cig_off, b->ix->shift,
                                                   b->d_hit_off64, b->d_hits, b->d_cigars,
                                                   (uint32_t)std::min<size_t>(b->hits_cap, 0xFFFFFFFFu), (uint32_t)std::min<size_t>(b->cigars_cap, 0xFFFFFFFFu),
                                                   b->sw.cig_cap, &b->d_ctr->abort);
    b->stats.launches += 1;
    K_CHECK(b->stream, "emit_kernel");
    return BWT_OK;
}

static int run_core(bwt_batch *b, bool with_sw, bool compact) {
    if (!b) { bwt_set_error("null batch"); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(b->ix->device));
    const float h2d = b->stats.ms_h2d;
    memset(&b->stats, 0, sizeof b->stats);
    b->stats.ms_h2d = h2d;
    b->sw.launches = 0;
    b->n_hits = 0; b->cigar_bytes = 0;
    b->has_result = false;
    const DevIndex &ix = b->ix->dv;
    const int64_t n = b->n;
    if (n == 0) {
        memset(b->h_ctr, 0, sizeof(SeedCounters));
        memset(b->sw.h_ctr, 0, sizeof(SwCounters));
        b->stats.reads = 0;
        b->has_result = with_sw; b->result_compact = compact;
        return BWT_OK;
    }
    for (int attempt = 0; attempt < 8; attempt++) {
        BWT_TRY(dev_grow(b->d_clusters_sparse, b->cl_sparse_cap, b->cl_cap, false));
        BWT_TRY(sw_stage_reserve(b->sw, (uint32_t)b->cl_cap));
        const uint32_t cl_cap = (uint32_t)std::min<size_t>(std::min(b->cl_cap, b->cl_sparse_cap), 0xFFFFFFF0u);
        b->stats.launches = 0; b->sw.launches = 0;
        CU_TRY(cudaMemsetAsync(b->d_totals, 0, sizeof(RunTotals), b->stream));
        CU_TRY(cudaEventRecord(b->ev[0], b->stream));
        if (b->need_pack) {           // ASCII batch: lengths, word offsets, 2-bit words (part of the work on such a batch)
            BWT_TRY(reads_geom_from_ascii(b->d_aoff, n, b->d_len, b->d_woff, b->d_scan_tmp, b->d_totals->info, b->stream, &b->stats.launches));
            BWT_TRY(reads_pack_device(b->d_ascii, b->d_aoff, b->R.aoff_base, n, b->d_len, b->d_woff, b->d_pk, b->d_exc, b->stream, &b->stats.launches));
        }
        BWT_TRY(enqueue_seed_stage(b, cl_cap));
        CU_TRY(cudaEventRecord(b->ev[1], b->stream));
        if (with_sw) {
            windows_kernel<<<148 * 8, 256, 0, b->stream>>>(b->d_clusters, &b->d_ctr->n_pairs, b->d_len, ix.ref_len_padded, b->sw.d_desc, b->sw.d_out);
            b->stats.launches += 1;
            K_CHECK(b->stream, "windows_kernel");
            b->sw.small = n <= (int64_t)std::max(0, sw_small_limit());
            BWT_TRY(sw_stage_enqueue(b->sw, b->R, ix.ref, ix.ref2, ix.refx, &b->d_ctr->n_pairs, cl_cap, b->max_read_len, b->stream));
            CU_TRY(cudaMemcpyAsync(b->sw.h_ctr, b->sw.d_ctr, sizeof(SwCounters), cudaMemcpyDeviceToHost, b->stream));
            CU_TRY(cudaEventRecord(b->ev[2], b->stream));
            BWT_TRY(enqueue_emit_stage(b, compact));
        } else {
            CU_TRY(cudaEventRecord(b->ev[2], b->stream));
        }
        CU_TRY(cudaEventRecord(b->ev[3], b->stream));
        CU_TRY(cudaMemcpyAsync(b->h_ctr, b->d_ctr, offsetof(SeedCounters, dense_base), cudaMemcpyDeviceToHost, b->stream));
        CU_TRY(cudaMemcpyAsync(b->h_totals, b->d_totals, sizeof(RunTotals), cudaMemcpyDeviceToHost, b->stream));
        CU_TRY(batch_wait(b, b->stream, b->ev_block, n > 65536));         // the one synchronisation of a run
        if (b->h_totals->info[1]) { bwt_set_error("offsets must be non-decreasing"); return BWT_ERR_ARG; }
        // capacity overflows of the seeding stage: grow and run again (rare: the capacities start generous)
        const SeedCounters &sc = *b->h_ctr;
        if (sc.n_slow > b->slow_cap) {
            cudaFree(b->d_slow); b->d_slow = nullptr;
            b->slow_cap = sc.n_slow + sc.n_slow / 4 + 1024;
            CU_TRY(cudaMalloc(&b->d_slow, (size_t)b->slow_cap * sizeof(SlowRead)));
            continue;
        }
        if (sc.scratch_overflow) {
            BWT_TRY(dev_grow(b->d_slow_scratch, b->slow_scratch_cap, (size_t)std::max<unsigned long long>(sc.scratch_used, 2 * b->slow_scratch_cap)));
            continue;
        }
        if (sc.cluster_overflow) {
            // a sub-list ran out of room: n_clusters is a lower bound of what is needed (refused reads not counted in full);
            // the regions are equal shares, so ask for twice the total, and every region must hold the fullest sub-list
            size_t want = (size_t)std::max<uint64_t>(sc.n_clusters, b->cl_cap) * 2 + 4096;
            std::vector<uint32_t> sub((size_t)SGA_SUBLISTS * 32);
            CU_TRY(cudaMemcpy(sub.data(), reinterpret_cast<const uint8_t *>(b->d_ctr) + offsetof(SeedCounters, sub_n),
                              sub.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            uint64_t fullest = 0;
            for (int r = 0; r < SGA_SUBLISTS; r++) fullest = std::max<uint64_t>(fullest, sub[(size_t)32 * r]);
            want = std::max<size_t>(want, (size_t)(fullest + fullest / 4 + 64) * SGA_SUBLISTS);
            if (want * sizeof(Cluster) > (48ull << 30)) { bwt_set_error("cluster list beyond 48 GiB: use smaller batches"); return BWT_ERR_NOMEM; }
            BWT_TRY(dev_grow(b->d_clusters, b->cl_cap, want, false));
            continue;
        }
        const uint32_t ncl = sc.n_clusters;
        if (with_sw) {
            bool redo_emit = false;
            const int rc = sw_stage_finish(b->sw, b->R, ix.ref, ix.ref2, ix.refx, cl_cap, b->max_read_len, b->stream, &redo_emit);
            if (rc == BWT_ERR_NOSPACE) continue;                         // CIGAR arena grown
            BWT_TRY(rc);
            if (redo_emit) {
                CU_TRY(cudaMemsetAsync(&b->d_totals->emit_flags, 0, sizeof(uint32_t), b->stream));
                BWT_TRY(enqueue_emit_stage(b, compact));
                CU_TRY(cudaEventRecord(b->ev[3], b->stream));
                CU_TRY(cudaMemcpyAsync(b->h_totals, b->d_totals, sizeof(RunTotals), cudaMemcpyDeviceToHost, b->stream));
                CU_TRY(cudaStreamSynchronize(b->stream));
            }
            if (compact && b->h_totals->emit_flags) { bwt_set_error("a hit does not fit the compact record (CIGAR beyond 65535 bytes): use the wide API"); return BWT_ERR_LIMIT; }
            b->n_hits = b->h_totals->hits;
            b->cigar_bytes = b->h_totals->cig_bytes;
        } else {
            b->n_hits = ncl;
        }
        bwt_stats &s = b->stats;
        s.reads = (uint64_t)n;
        s.searches = sc.searches; s.lf_steps = sc.lf_steps; s.hbm_lf_steps = sc.hbm_lf_steps;
        s.located_rows = sc.located;
        s.ms_sw_kernel = with_sw ? b->sw.ms_fast_kernel : 0.f;
        s.clusters = ncl;
        s.sw_calls = b->sw.sw_calls; s.sw_cells = b->sw.cells; s.sw_general = b->sw.n_general;
        s.sw_shortcut = b->sw.shortcut_calls; s.sw_shortcut_cells = b->sw.shortcut_cells;
        s.sw_band = b->sw.band_calls; s.sw_band_cells = b->sw.band_cells; s.sw_band_filled = b->sw.band_cells_filled;
        s.launches += b->sw.launches;
        cudaEventElapsedTime(&s.ms_sga_kernel, b->ev[6], b->ev[7]);
        cudaEventElapsedTime(&s.ms_seed, b->ev[0], b->ev[1]);
        cudaEventElapsedTime(&s.ms_sw, b->ev[1], b->ev[2]);
        cudaEventElapsedTime(&s.ms_emit, b->ev[2], b->ev[3]);
        cudaEventElapsedTime(&s.ms_total, b->ev[0], b->ev[3]);
        b->has_result = with_sw; b->result_compact = compact;
        return BWT_OK;
    }
    bwt_set_error("a capacity kept overflowing");
    return BWT_ERR_NOMEM;
}

extern "C" int bwt_batch_run(bwt_batch *b) { return run_core(b, true, false); }
extern "C" int bwt_batch_run_compact(bwt_batch *b) { return run_core(b, true, true); }

extern "C" int bwt_batch_result_size(bwt_batch *b, int64_t *n_hits, int64_t *cigar_bytes) {
    if (!b) { bwt_set_error("null batch"); return BWT_ERR_ARG; }
    if (n_hits) *n_hits = b->n_hits;
    if (cigar_bytes) *cigar_bytes = b->cigar_bytes;
    return BWT_OK;
}

extern "C" int bwt_batch_download(bwt_batch *b, int32_t *read_status, int64_t *hit_offsets, bwt_hit *hits,
                                  int64_t hits_cap, char *cigars, int64_t cigar_cap) {
    if (!b || !b->has_result || b->result_compact) { bwt_set_error("bwt_batch_download: no result (call bwt_batch_run first)"); return BWT_ERR_ARG; }
    if (hits_cap < b->n_hits || cigar_cap < b->cigar_bytes) {
        bwt_set_error("output capacity too small: need %lld hits and %lld CIGAR bytes", (long long)b->n_hits, (long long)b->cigar_bytes);
        return BWT_ERR_NOSPACE;
    }
    CU_TRY(cudaSetDevice(b->ix->device));
    CU_TRY(cudaEventRecord(b->ev[8], b->stream));
    if (read_status && b->n) CU_TRY(cudaMemcpyAsync(read_status, b->d_read_status, (size_t)b->n * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    if (hit_offsets) {
        if (b->n) CU_TRY(cudaMemcpyAsync(hit_offsets, b->d_hit_off64, (size_t)(b->n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, b->stream));
        else hit_offsets[0] = 0;
    }
    if (hits && b->n_hits) CU_TRY(cudaMemcpyAsync(hits, b->d_hits, (size_t)b->n_hits * sizeof(bwt_hit), cudaMemcpyDeviceToHost, b->stream));
    if (cigars && b->cigar_bytes) CU_TRY(cudaMemcpyAsync(cigars, b->d_cigars, (size_t)b->cigar_bytes, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaEventRecord(b->ev[9], b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    cudaEventElapsedTime(&b->stats.ms_d2h, b->ev[8], b->ev[9]);
    return BWT_OK;
}

extern "C" int bwt_batch_download_compact(bwt_batch *b, uint8_t *read_status, uint32_t *hit_offsets, bwt_hit16 *hits,
                                          int64_t hits_cap, char *cigars, int64_t cigar_cap) {
    if (!b || !b->has_result || !b->result_compact) { bwt_set_error("bwt_batch_download_compact: no compact result (call bwt_batch_run_compact first)"); return BWT_ERR_ARG; }
    if (hits_cap < b->n_hits || cigar_cap < b->cigar_bytes) {
        bwt_set_error("output capacity too small: need %lld hits and %lld CIGAR bytes", (long long)b->n_hits, (long long)b->cigar_bytes);
        return BWT_ERR_NOSPACE;
    }
    CU_TRY(cudaSetDevice(b->ix->device));
    CU_TRY(cudaEventRecord(b->ev[8], b->stream));
    if (read_status && b->n) CU_TRY(cudaMemcpyAsync(read_status, b->d_status8, (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
    if (hit_offsets) {
        if (b->n) CU_TRY(cudaMemcpyAsync(hit_offsets, b->d_hit_off, (size_t)(b->n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
        else hit_offsets[0] = 0;
    }
    if (hits && b->n_hits) CU_TRY(cudaMemcpyAsync(hits, b->d_hits16, (size_t)b->n_hits * sizeof(bwt_hit16), cudaMemcpyDeviceToHost, b->stream));
    if (cigars && b->cigar_bytes) CU_TRY(cudaMemcpyAsync(cigars, b->d_cigars, (size_t)b->cigar_bytes, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaEventRecord(b->ev[9], b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    cudaEventElapsedTime(&b->stats.ms_d2h, b->ev[8], b->ev[9]);
    return BWT_OK;
}

extern "C" int bwt_batch_stats(bwt_batch *b, bwt_stats *out) {
    if (!b || !out) { bwt_set_error("bwt_batch_stats: null argument"); return BWT_ERR_ARG; }
    *out = b->stats;
    return BWT_OK;
}

extern "C" int bwt_batch_device_result(bwt_batch *b, void **d_hit_offsets, void **d_hits, void **d_cigars) {
    if (!b || !b->has_result) { bwt_set_error("bwt_batch_device_result: no result"); return BWT_ERR_ARG; }
    if (d_hit_offsets) *d_hit_offsets = b->result_compact ? (void *)b->d_hit_off : (void *)b->d_hit_off64;
    if (d_hits) *d_hits = b->result_compact ? (void *)b->d_hits16 : (void *)b->d_hits;
    if (d_cigars) *d_cigars = b->d_cigars;
    return BWT_OK;
}

// internals the partition gather (gather.cu) needs
int batch_result_view(bwt_batch *b, int *device, cudaStream_t *stream, int64_t *n, const uint8_t **d_status8, const uint32_t **d_hit_off,
                      const bwt_hit16 **d_hits16, const char **d_cigars, int64_t *n_hits, int64_t *cigar_bytes) {
    if (!b || !b->has_result || !b->result_compact) { bwt_set_error("partition gather needs a compact result (bwt_batch_run_compact)"); return BWT_ERR_ARG; }
    *device = b->ix->device; *stream = b->stream; *n = b->n; *d_status8 = b->d_status8; *d_hit_off = b->d_hit_off; *d_hits16 = b->d_hits16;
    *d_cigars = b->d_cigars; *n_hits = b->n_hits; *cigar_bytes = b->cigar_bytes;
    return BWT_OK;
}

// ---- one-shot calls ---------------------------------------------------------------------------------

static const int64_t CHUNK_READS = 1 << 20;
static const int64_t CHUNK_BYTES = 256ll << 20;

// splits [0,n) into chunks that respect both limits; returns the end of the chunk starting at lo
static int64_t chunk_end_lim(const int64_t *offsets, int64_t lo, int64_t n, int64_t max_reads, int64_t max_bytes) {
    int64_t hi = std::min(n, lo + max_reads);
    if (offsets[hi] - offsets[lo] > max_bytes) {           // last index whose prefix still fits
        const int64_t *e = std::upper_bound(offsets + lo, offsets + hi + 1, offsets[lo] + max_bytes);
        hi = std::max<int64_t>(lo + 1, (e - offsets) - 1);
    }
    return hi;
}
static int64_t chunk_end(const int64_t *offsets, int64_t lo, int64_t n) { return chunk_end_lim(offsets, lo, n, CHUNK_READS, CHUNK_BYTES); }

// ---- pipelined aligner ------------------------------------------------------------------------------
// `depth` resident batches, one host thread each; chunk c is uploaded, run and downloaded by thread c mod depth on
// that batch's own stream, so the H2D of one chunk, the kernels of another and the D2H of a third overlap -- and so do
// the kernels of two chunks (the seeding kernel waits on HBM, the extension kernels on the ALU pipe).  Results land in
// the caller's arrays at their final place: a chunk learns its bases once every earlier chunk has finished.
struct bwt_aligner {
    bwt_index *ix = nullptr;
    int64_t chunk_reads = 0, chunk_bytes = 0;
    std::vector<bwt_batch *> batches;
    bwt_stats stats{};
};

extern "C" int bwt_aligner_create(bwt_index *h, int64_t chunk_reads, int64_t chunk_bytes, int depth, bwt_aligner **out) {
    if (!h || !out || chunk_reads <= 0 || chunk_bytes <= 0 || depth < 1 || depth > 16) { bwt_set_error("bwt_aligner_create: bad argument"); return BWT_ERR_ARG; }
    bwt_aligner *a = new bwt_aligner();
    a->ix = h; a->chunk_reads = chunk_reads; a->chunk_bytes = chunk_bytes;
    for (int k = 0; k < depth; k++) {
        bwt_batch *b = nullptr;
        int st = bwt_batch_create(h, chunk_reads, chunk_bytes, &b);
        if (st != BWT_OK) { bwt_aligner_destroy(a); return st; }
        a->batches.push_back(b);
    }
    *out = a;
    return BWT_OK;
}

extern "C" int bwt_aligner_destroy(bwt_aligner *a) {
    if (!a) return BWT_OK;
    for (bwt_batch *b : a->batches) bwt_batch_destroy(b);
    delete a;
    return BWT_OK;
}

extern "C" int bwt_aligner_stats(bwt_aligner *a, bwt_stats *out) {
    if (!a || !out) { bwt_set_error("bwt_aligner_stats: null argument"); return BWT_ERR_ARG; }
    *out = a->stats;
    return BWT_OK;
}

namespace {
struct PipeShared {
    std::mutex mu, compute_mu;
    std::condition_variable cv, compute_cv;
    int64_t compute_next = 0;        // next chunk to enter its run (chunks enter in order)
    int compute_in = 0;              // chunks inside their run
    int64_t next_chunk = 0;          // chunks [0, next_chunk) have published their totals
    int64_t hit_base = 0, cig_base = 0;
    bool space = true;
    int status = BWT_OK;
    std::string error;
};

void stats_add(bwt_stats &t, const bwt_stats &s) {
    t.reads += s.reads; t.searches += s.searches; t.lf_steps += s.lf_steps; t.hbm_lf_steps += s.hbm_lf_steps;
    t.located_rows += s.located_rows; t.clusters += s.clusters; t.sw_calls += s.sw_calls; t.sw_cells += s.sw_cells;
    t.sw_shortcut += s.sw_shortcut; t.sw_shortcut_cells += s.sw_shortcut_cells;
    t.sw_band += s.sw_band; t.sw_band_cells += s.sw_band_cells; t.sw_band_filled += s.sw_band_filled;
    t.sw_general += s.sw_general; t.ms_total += s.ms_total; t.ms_seed += s.ms_seed; t.ms_sw += s.ms_sw; t.ms_emit += s.ms_emit;
    t.ms_h2d += s.ms_h2d; t.ms_d2h += s.ms_d2h; t.ms_sga_kernel += s.ms_sga_kernel; t.ms_sw_kernel += s.ms_sw_kernel;
    t.launches += s.launches;
}

// the two faces of an aligner run: ASCII in / wide records out (the reference-shaped call), packed in / compact out
struct AlignIO {
    // input
    const uint8_t *reads = nullptr; const int64_t *offsets = nullptr;     // ASCII
    const bwt_packed_reads *packed = nullptr;                             // packed
    int64_t n = 0;
    // output
    bool compact = false;
    int32_t *status32 = nullptr; int64_t *off64 = nullptr; bwt_hit *hits = nullptr;
    uint8_t *status8 = nullptr; uint32_t *off32 = nullptr; bwt_hit16 *hits16 = nullptr;
    int64_t hits_cap = 0; char *cigars = nullptr; int64_t cigar_cap = 0;
    int64_t *n_hits = nullptr, *cigar_bytes = nullptr;
};

int aligner_run_io(bwt_aligner *a, const AlignIO &io) {
    memset(&a->stats, 0, sizeof a->stats);
    if (io.compact) io.off32[0] = 0; else io.off64[0] = 0;
    *io.n_hits = 0; *io.cigar_bytes = 0;
    const int64_t n = io.n;
    if (n == 0) return BWT_OK;
    std::vector<int64_t> bounds(1, 0);                      // chunk c = reads [bounds[c], bounds[c+1])
    std::vector<int64_t> wbase, wcount;                     // packed input: word base / count / longest read per chunk
    std::vector<int> wmax;
    if (io.packed) {
        int64_t wb = 0;
        for (int64_t lo = 0; lo < n;) {
            const int64_t hi = std::min(n, lo + a->chunk_reads);
            int64_t words; int mx;
            packed_geom(io.packed, lo, hi, &words, &mx);
            wbase.push_back(wb); wcount.push_back(words); wmax.push_back(mx);
            wb += words;
            bounds.push_back(hi);
            lo = hi;
        }
    } else {
        for (int64_t lo = 0; lo < n;) {
            int64_t hi = chunk_end_lim(io.offsets, lo, n, a->chunk_reads, a->chunk_bytes);
            if (io.offsets[hi] - io.offsets[lo] > a->chunk_bytes) { bwt_set_error("a single read exceeds %lld bytes", (long long)a->chunk_bytes); return BWT_ERR_LIMIT; }
            bounds.push_back(hi);
            lo = hi;
        }
    }
    const int64_t nchunks = (int64_t)bounds.size() - 1;
    const int depth = (int)std::min<int64_t>((int64_t)a->batches.size(), nchunks);
    PipeShared sh;
    std::vector<bwt_stats> wstats((size_t)depth);
    const bool trace = getenv("BWT_TRACE") != nullptr;       // per-chunk host timeline on stderr (debug aid)
    const char *env_tok = getenv("BWT_COMPUTE_TOKEN");
    // Chunks inside their run (kernels enqueued or executing) at any time, entered in chunk order.  With free overlap the
    // workers fall into lock step -- all `depth` chunks compute side by side, finish together, then all wait for their
    // copies while the GPU idles (profiles/r2b3_e2e_timeline.md: 1.3 ms per round of four).  Three at a time keeps the next
    // chunk's kernels queued behind the running ones' while the other workers' copies go on beside them (sweeps in
    // profiles/r2b4_e2e_token_sweep.log, r2b6: three at a time with five chunks in flight).
    // BWT_COMPUTE_TOKEN=0: free overlap; N > 0: at most N.
    const int token_mode = env_tok && *env_tok ? atoi(env_tok) : 3;
    cudaEvent_t ev_origin = nullptr;                         // trace: device-side times of every chunk against one origin
    if (trace) { cudaSetDevice(a->batches[0]->ix->device); cudaEventCreate(&ev_origin); cudaEventRecord(ev_origin, a->batches[0]->stream); }
    auto dev_ms = [&](cudaEvent_t e) { float ms = -1.f; if (ev_origin) cudaEventElapsedTime(&ms, ev_origin, e); return ms; };
    const auto t_start = std::chrono::steady_clock::now();
    auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    auto worker = [&](int w) {
        bwt_batch *b = a->batches[(size_t)w];
        memset(&wstats[(size_t)w], 0, sizeof(bwt_stats));
        for (int64_t c = w; c < nchunks; c += depth) {
            const int64_t lo = bounds[(size_t)c], m = bounds[(size_t)c + 1] - lo;
            int st = BWT_OK;
            bool skip;
            { std::lock_guard<std::mutex> g(sh.mu); skip = sh.status != BWT_OK; }
            double t0 = now_ms(), t1 = t0, t2 = t0, t3 = t0;
            bool took_turn = false;
            if (!skip) {
                if (b->dl_pending) {     // the previous chunk's results must have left before this run overwrites them
                    cudaError_t e = batch_wait(b, b->dl_stream, b->ev_block_dl, b->n > 65536);
                    b->dl_pending = false;
                    if (e != cudaSuccess) { bwt_set_error("aligner download: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; }
                    else {
                        float ms = 0; cudaEventElapsedTime(&ms, b->ev_dl0, b->ev_dl1); wstats[(size_t)w].ms_d2h += ms;
                        if (trace) fprintf(stderr, "chunk %3lld w%d: device d2h %.2f-%.2f (collected at host %.2f)\n", (long long)(c - depth), w, dev_ms(b->ev_dl0), dev_ms(b->ev_dl1), now_ms());
                    }
                }
                if (st == BWT_OK)
                    st = io.packed ? batch_upload_packed_range(b, io.packed, lo, lo + m, wbase[(size_t)c], wcount[(size_t)c], wmax[(size_t)c])
                                   : bwt_batch_upload(b, io.reads, io.offsets + lo, m);
                t1 = now_ms();
                if (st == BWT_OK) {
                    if (token_mode > 0) {
                        std::unique_lock<std::mutex> g(sh.compute_mu);
                        sh.compute_cv.wait(g, [&] { return sh.compute_next == c && sh.compute_in < token_mode; });
                        sh.compute_next = c + 1; sh.compute_in++;
                        took_turn = true;
                    }
                    sh.compute_cv.notify_all();
                    st = run_core(b, true, io.compact);
                    if (token_mode > 0) { std::lock_guard<std::mutex> g(sh.compute_mu); sh.compute_in--; }
                    sh.compute_cv.notify_all();
                    if (st == BWT_OK) { float ms = 0; cudaEventElapsedTime(&ms, b->ev[4], b->ev[5]); b->stats.ms_h2d = ms; }
                }
                t2 = now_ms();
            }
            if (token_mode > 0 && !took_turn) {
                // a chunk that does not run (an earlier failure, a failed upload) still takes its turn: the chunks behind it wait for it
                std::unique_lock<std::mutex> g(sh.compute_mu);
                sh.compute_cv.wait(g, [&] { return sh.compute_next == c; });
                sh.compute_next = c + 1;
                g.unlock();
                sh.compute_cv.notify_all();
            }
            int64_t nh = (skip || st != BWT_OK) ? 0 : b->n_hits, nc = (skip || st != BWT_OK) ? 0 : b->cigar_bytes;
            int64_t hb, cb;
            bool room;
            {
                std::unique_lock<std::mutex> g(sh.mu);
                if (st != BWT_OK && sh.status == BWT_OK) { sh.status = st; sh.error = bwt_last_error(); }
                sh.cv.wait(g, [&] { return sh.next_chunk == c; });
                hb = sh.hit_base; cb = sh.cig_base;
                sh.hit_base += nh; sh.cig_base += nc;
                if (!((io.compact ? (void *)io.hits16 : (void *)io.hits) && io.cigars && sh.hit_base <= io.hits_cap && sh.cig_base <= io.cigar_cap)) sh.space = false;
                if (io.compact && (sh.hit_base >= (1ll << 32) || sh.cig_base >= (1ll << 32)) && sh.status == BWT_OK) {
                    sh.status = BWT_ERR_LIMIT; sh.error = "compact results address 2^32 hits / CIGAR bytes per run: split the call";
                }
                room = sh.space && sh.status == BWT_OK;
                sh.next_chunk = c + 1;
            }
            sh.cv.notify_all();
            t3 = now_ms();
            if (skip || st != BWT_OK) continue;
            if (room) {
                const int64_t cnt = std::max(m + 1, nh);
                if (io.compact) rebase_result16_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, b->stream>>>(b->d_hit_off, m, b->d_hits16, nh, (uint32_t)hb, (uint32_t)cb);
                else rebase_result_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, b->stream>>>(b->d_hit_off64, m, b->d_hits, nh, hb, cb);
                // entry m of this chunk equals entry 0 of the next one: only the last chunk writes it
                st = cudaGetLastError() == cudaSuccess ? BWT_OK : BWT_ERR_CUDA;
                if (st == BWT_OK) {
                    // D2H on the batch's download stream, behind the rebase kernel; not waited for here: the worker goes on
                    // to upload its next chunk (other buffers) and collects the copy before its next run
                    cudaEventRecord(b->ev_run_done, b->stream);
                    cudaStream_t s = b->dl_stream;
                    cudaStreamWaitEvent(s, b->ev_run_done, 0);
                    cudaEventRecord(b->ev_dl0, s);
                    const size_t last = c + 1 == nchunks ? 1 : 0;
                    if (io.compact) {
                        if (io.status8) cudaMemcpyAsync(io.status8 + lo, b->d_status8, (size_t)m, cudaMemcpyDeviceToHost, s);
                        cudaMemcpyAsync(io.off32 + lo, b->d_hit_off, ((size_t)m + last) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
                        if (nh) cudaMemcpyAsync(io.hits16 + hb, b->d_hits16, (size_t)nh * sizeof(bwt_hit16), cudaMemcpyDeviceToHost, s);
                    } else {
                        if (io.status32) cudaMemcpyAsync(io.status32 + lo, b->d_read_status, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
                        cudaMemcpyAsync(io.off64 + lo, b->d_hit_off64, ((size_t)m + last) * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
                        if (nh) cudaMemcpyAsync(io.hits + hb, b->d_hits, (size_t)nh * sizeof(bwt_hit), cudaMemcpyDeviceToHost, s);
                    }
                    if (nc) cudaMemcpyAsync(io.cigars + cb, b->d_cigars, (size_t)nc, cudaMemcpyDeviceToHost, s);
                    cudaEventRecord(b->ev_dl1, s);
                    if (cudaGetLastError() != cudaSuccess) { bwt_set_error("aligner download: enqueue failed"); st = BWT_ERR_CUDA; }
                    else b->dl_pending = true;
                }
                if (st != BWT_OK) { std::lock_guard<std::mutex> g(sh.mu); if (sh.status == BWT_OK) { sh.status = st; sh.error = bwt_last_error(); } }
            }
            stats_add(wstats[(size_t)w], b->stats);
            if (trace)
                fprintf(stderr, "chunk %3lld w%d: host upload %.2f-%.2f run -%.2f wait -%.2f download -%.2f | device h2d %.2f-%.2f seed %.2f sw %.2f emit %.2f-%.2f\n", (long long)c, w,
                        t0, t1, t2, t3, now_ms(), dev_ms(b->ev[4]), dev_ms(b->ev[5]), dev_ms(b->ev[0]), dev_ms(b->ev[1]), dev_ms(b->ev[2]), dev_ms(b->ev[3]));
        }
    };
    auto worker_all = [&](int w) {
        worker(w);
        bwt_batch *b = a->batches[(size_t)w];
        if (b->dl_pending) {
            cudaError_t e = batch_wait(b, b->dl_stream, b->ev_block_dl, b->n > 65536);
            b->dl_pending = false;
            if (e != cudaSuccess) { std::lock_guard<std::mutex> g(sh.mu); if (sh.status == BWT_OK) { sh.status = BWT_ERR_CUDA; sh.error = cudaGetErrorString(e); } }
            else {
                float ms = 0; cudaEventElapsedTime(&ms, b->ev_dl0, b->ev_dl1); wstats[(size_t)w].ms_d2h += ms;
                if (trace) fprintf(stderr, "last of   w%d: device d2h %.2f-%.2f (collected at host %.2f)\n", w, dev_ms(b->ev_dl0), dev_ms(b->ev_dl1), now_ms());
            }
        }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < depth; w++) th.emplace_back(worker_all, w);
    worker_all(0);
    for (auto &t : th) t.join();
    if (ev_origin) cudaEventDestroy(ev_origin);
    for (int w = 0; w < depth; w++) stats_add(a->stats, wstats[(size_t)w]);
    *io.n_hits = sh.hit_base; *io.cigar_bytes = sh.cig_base;
    if (sh.status != BWT_OK) { bwt_set_error("%s", sh.error.c_str()); return sh.status; }
    if (!sh.space) {
        bwt_set_error("output capacity too small: need %lld hits and %lld CIGAR bytes", (long long)sh.hit_base, (long long)sh.cig_base);
        return BWT_ERR_NOSPACE;
    }
    return BWT_OK;
}
}  // namespace

extern "C" int bwt_aligner_run(bwt_aligner *a, const uint8_t *reads, const int64_t *offsets, int64_t n, int32_t *read_status,
                               int64_t *hit_offsets, bwt_hit *hits, int64_t hits_cap, int64_t *n_hits, char *cigars,
                               int64_t cigar_cap, int64_t *cigar_bytes) {
    if (!a || !offsets || n < 0 || !hit_offsets || !n_hits || !cigar_bytes || (n > 0 && !reads)) { bwt_set_error("bwt_aligner_run: null argument"); return BWT_ERR_ARG; }
    AlignIO io;
    io.reads = reads; io.offsets = offsets; io.n = n;
    io.status32 = read_status; io.off64 = hit_offsets; io.hits = hits; io.hits_cap = hits_cap; io.cigars = cigars; io.cigar_cap = cigar_cap;
    io.n_hits = n_hits; io.cigar_bytes = cigar_bytes;
    return aligner_run_io(a, io);
}

extern "C" int bwt_aligner_run_packed(bwt_aligner *a, const bwt_packed_reads *p, uint8_t *read_status, uint32_t *hit_offsets,
                                      bwt_hit16 *hits, int64_t hits_cap, int64_t *n_hits, char *cigars, int64_t cigar_cap,
                                      int64_t *cigar_bytes) {
    if (!a || !hit_offsets || !n_hits || !cigar_bytes) { bwt_set_error("bwt_aligner_run_packed: null argument"); return BWT_ERR_ARG; }
    BWT_TRY(packed_check(p));
    AlignIO io;
    io.packed = p; io.n = p->n; io.compact = true;
    io.status8 = read_status; io.off32 = hit_offsets; io.hits16 = hits; io.hits_cap = hits_cap; io.cigars = cigars; io.cigar_cap = cigar_cap;
    io.n_hits = n_hits; io.cigar_bytes = cigar_bytes;
    return aligner_run_io(a, io);
}

// worker_bwt.erl:73-97 for one list of reads, host buffers in and out: a short-lived aligner
void aligner_cache_drop(bwt_index *h) {
    if (!h || !h->cached_aligner) return;
    bwt_aligner_destroy(static_cast<bwt_aligner *>(h->cached_aligner));
    h->cached_aligner = nullptr;
    h->cached_busy = false;
}

static int64_t round_up_pow2(int64_t v, int64_t lo) {
    int64_t p = lo;
    while (p < v) p <<= 1;
    return p;
}

extern "C" int bwt_align_batch(bwt_index *h, const uint8_t *reads, const int64_t *offsets, int64_t n,
                               int32_t *read_status, int64_t *hit_offsets, bwt_hit *hits, int64_t hits_cap,
                               int64_t *n_hits, char *cigars, int64_t cigar_cap, int64_t *cigar_bytes) {
    if (!h || !offsets || n < 0 || !hit_offsets || !n_hits || !cigar_bytes) { bwt_set_error("bwt_align_batch: null argument"); return BWT_ERR_ARG; }
    const int64_t total_bytes = offsets[n] - offsets[0];
    // sizes rounded up to powers of two: the aligner kept with the index then fits the next call as well
    const int64_t cr = std::min(round_up_pow2(std::max<int64_t>(1, n), 1024), CHUNK_READS);
    const int64_t cb = std::min(round_up_pow2(std::max<int64_t>(16, total_bytes), 1 << 17), CHUNK_BYTES);
    const int depth = (int)std::min<int64_t>(5, std::max<int64_t>(1, (n + CHUNK_READS - 1) / CHUNK_READS));   // chunks in flight (profiles/r2b6: five, three of them inside their run)
    bwt_aligner *a = nullptr;
    bool temporary = false;
    if (h->tail_mu) {
        std::lock_guard<std::mutex> g(*h->tail_mu);
        if (h->cached_aligner && !h->cached_busy && h->cached_reads >= cr && h->cached_bytes >= cb && h->cached_depth >= depth) {
            a = static_cast<bwt_aligner *>(h->cached_aligner);
            h->cached_busy = true;
        }
    }
    if (!a) {
        BWT_TRY(bwt_aligner_create(h, cr, cb, depth, &a));
        temporary = true;
        if (h->tail_mu) {
            std::lock_guard<std::mutex> g(*h->tail_mu);
            if (!h->cached_busy) {                          // keep the new one (it is at least as large in the dimension that missed)
                if (h->cached_aligner) bwt_aligner_destroy(static_cast<bwt_aligner *>(h->cached_aligner));
                h->cached_aligner = a; h->cached_reads = cr; h->cached_bytes = cb; h->cached_depth = depth;
                h->cached_busy = true;
                temporary = false;
            }
        }
    }
    int st = bwt_aligner_run(a, reads, offsets, n, read_status, hit_offsets, hits, hits_cap, n_hits, cigars, cigar_cap, cigar_bytes);
    if (temporary) bwt_aligner_destroy(a);
    else { std::lock_guard<std::mutex> g(*h->tail_mu); h->cached_busy = false; }
    return st;
}

extern "C" int bwt_sga_batch(bwt_index *h, const uint8_t *reads, const int64_t *offsets, int64_t n,
                             int32_t *read_status, int64_t *cl_offsets, int64_t *S, int32_t *D, int64_t cap,
                             int64_t *total) {
    if (!h || !offsets || n < 0 || !cl_offsets || !total) { bwt_set_error("bwt_sga_batch: null argument"); return BWT_ERR_ARG; }
    int64_t total_bytes = offsets[n] - offsets[0];
    bwt_batch *b = nullptr;
    BWT_TRY(bwt_batch_create(h, std::max<int64_t>(1, std::min(n, CHUNK_READS)), std::max<int64_t>(16, std::min(total_bytes, CHUNK_BYTES)), &b));
    int st = BWT_OK;
    int64_t base = 0;
    bool space = true;
    cl_offsets[0] = 0;
    int64_t *d_S = nullptr;
    int32_t *d_D = nullptr;
    size_t capS = 0, capD = 0;
    for (int64_t lo = 0; lo < n && st == BWT_OK;) {
        int64_t hi = chunk_end(offsets, lo, n), m = hi - lo;
        if ((st = bwt_batch_upload(b, reads, offsets + lo, m)) != BWT_OK) break;
        if ((st = run_core(b, false, false)) != BWT_OK) break;
        int64_t nc = b->n_hits;
        // offsets = exclusive scan of the per-read cluster counts
        cudaMemcpyAsync(b->d_hit_off, b->d_ncl, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, b->stream);
        cudaMemsetAsync(b->d_hit_off + m, 0, sizeof(uint32_t), b->stream);
        if ((st = exclusive_scan_u32(b->d_hit_off, (size_t)m + 1, b->d_scan_tmp, nullptr, b->stream, nullptr)) != BWT_OK) break;
        if ((st = dev_grow(d_S, capS, (size_t)nc + 1)) != BWT_OK || (st = dev_grow(d_D, capD, (size_t)nc + 1)) != BWT_OK) break;
        emit_sga_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, b->stream>>>(b->d_ncl, b->d_cl_base, b->d_clusters, m,
                                                                               b->d_hit_off, b->d_hit_off64, d_S, d_D);
        cudaStreamSynchronize(b->stream);
        if (cudaGetLastError() != cudaSuccess) { bwt_set_error("emit_sga_kernel failed"); st = BWT_ERR_CUDA; break; }
        cudaMemcpy(cl_offsets + lo, b->d_hit_off64, (size_t)(m + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost);
        if (base) for (int64_t r = lo; r <= hi; r++) cl_offsets[r] += base;
        if (read_status) cudaMemcpy(read_status + lo, b->d_read_status, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (space && S && D && base + nc <= cap) {
            if (nc) {
                cudaMemcpy(S + base, d_S, (size_t)nc * sizeof(int64_t), cudaMemcpyDeviceToHost);
                cudaMemcpy(D + base, d_D, (size_t)nc * sizeof(int32_t), cudaMemcpyDeviceToHost);
            }
        } else space = false;
        base += nc;
        lo = hi;
    }
    cudaFree(d_S); cudaFree(d_D);
    bwt_batch_destroy(b);
    if (st != BWT_OK) return st;
    *total = base;
    if (!space) { bwt_set_error("output capacity too small: need %lld clusters", (long long)base); return BWT_ERR_NOSPACE; }
    return BWT_OK;
}

extern "C" int bwt_search_batch(bwt_index *h, const uint8_t *subs, const int64_t *offsets, int64_t n,
                                int32_t *status, uint32_t *sp, uint32_t *ep, int32_t *nmatched) {
    if (!h || !offsets || n < 0 || (n > 0 && !subs)) { bwt_set_error("bwt_search_batch: null argument"); return BWT_ERR_ARG; }
    if (n == 0) return BWT_OK;
    BWT_TRY(index_select_device(h->device));
    int64_t base = offsets[0], bytes = offsets[n] - base;
    uint8_t *d_subs = nullptr;
    int64_t *d_offs = nullptr;
    int32_t *d_status = nullptr, *d_n = nullptr;
    uint32_t *d_sp = nullptr, *d_ep = nullptr;
    int st = BWT_OK;
    std::vector<int64_t> rel((size_t)n + 1);
    for (int64_t k = 0; k <= n; k++) rel[k] = offsets[k] - base;
    do {
        if (cudaMalloc(&d_subs, (size_t)bytes + 16) != cudaSuccess || cudaMalloc(&d_offs, (size_t)(n + 1) * 8) != cudaSuccess ||
            cudaMalloc(&d_status, (size_t)n * 4) != cudaSuccess || cudaMalloc(&d_n, (size_t)n * 4) != cudaSuccess ||
            cudaMalloc(&d_sp, (size_t)n * 4) != cudaSuccess || cudaMalloc(&d_ep, (size_t)n * 4) != cudaSuccess) {
            bwt_set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_NOMEM; break;
        }
        cudaMemcpy(d_subs, subs + base, (size_t)bytes, cudaMemcpyHostToDevice);
        cudaMemcpy(d_offs, rel.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
        st = launch_search(h->dv, d_subs, d_offs, n, 0, d_status, d_sp, d_ep, d_n, nullptr, 0);
        if (st != BWT_OK) break;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { bwt_set_error("search kernel failed: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; }
        if (status) cudaMemcpy(status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if (sp) cudaMemcpy(sp, d_sp, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if (ep) cudaMemcpy(ep, d_ep, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if (nmatched) cudaMemcpy(nmatched, d_n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_subs); cudaFree(d_offs); cudaFree(d_status); cudaFree(d_n); cudaFree(d_sp); cudaFree(d_ep);
    return st;
}

extern "C" int bwt_locate(bwt_index *h, uint32_t first_row, uint32_t count, uint32_t *sa_out) {
    if (!h || !sa_out) { bwt_set_error("bwt_locate: null argument"); return BWT_ERR_ARG; }
    if (first_row < 1 || (uint64_t)first_row + count - 1 > h->n) { bwt_set_error("bwt_locate: rows %u..+%u outside 1..%llu", first_row, count, (unsigned long long)h->n); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaMemcpy(sa_out, (const uint32_t *)h->d_sa + (first_row - 1), (size_t)count * 4, cudaMemcpyDeviceToHost));
    return BWT_OK;
}

extern "C" int bwt_search_resident(bwt_index *h, const uint8_t *subs, int64_t n, int32_t sub_len, int32_t repeats,
                                   float *ms, uint64_t *lf_steps, uint64_t *hbm_lf_steps) {
    if (!h || !subs || n <= 0 || sub_len <= 0 || repeats <= 0) { bwt_set_error("bwt_search_resident: bad argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(h->device));
    uint8_t *d_subs = nullptr;
    uint32_t *d_sp = nullptr, *d_ep = nullptr;
    SeedCounters *d_ctr = nullptr;
    cudaEvent_t e0, e1;
    CU_TRY(cudaMalloc(&d_subs, (size_t)n * sub_len + 16));
    CU_TRY(cudaMalloc(&d_sp, (size_t)n * 4));
    CU_TRY(cudaMalloc(&d_ep, (size_t)n * 4));
    CU_TRY(cudaMalloc(&d_ctr, sizeof(SeedCounters)));
    CU_TRY(cudaMemcpy(d_subs, subs, (size_t)n * sub_len, cudaMemcpyHostToDevice));
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    // warm-up launch, then the timed repeats
    BWT_TRY(launch_search(h->dv, d_subs, nullptr, n, sub_len, nullptr, d_sp, d_ep, nullptr, nullptr, 0));
    CU_TRY(cudaMemset(d_ctr, 0, sizeof(SeedCounters)));
    CU_TRY(cudaEventRecord(e0, 0));
    for (int r = 0; r < repeats; r++)
        BWT_TRY(launch_search(h->dv, d_subs, nullptr, n, sub_len, nullptr, d_sp, d_ep, nullptr, d_ctr, 0));
    CU_TRY(cudaEventRecord(e1, 0));
    CU_TRY(cudaEventSynchronize(e1));
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    SeedCounters c;
    CU_TRY(cudaMemcpy(&c, d_ctr, sizeof c, cudaMemcpyDeviceToHost));
    if (ms) *ms = t;
    if (lf_steps) *lf_steps = c.lf_steps / repeats;
    if (hbm_lf_steps) *hbm_lf_steps = c.hbm_lf_steps / repeats;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_subs); cudaFree(d_sp); cudaFree(d_ep); cudaFree(d_ctr);
    return BWT_OK;
}

extern "C" int bwt_sw_batch(int device, const uint8_t *q, const int64_t *q_offsets, const uint8_t *r,
                            const int64_t *r_offsets, int64_t n, int32_t *status, int32_t *score, int64_t *cigar_off,
                            int32_t *cigar_len, char *cigars, int64_t cigar_cap, int64_t *cigar_total) {
    if (!q_offsets || !r_offsets || n < 0 || !status || !score || !cigar_off || !cigar_len || !cigar_total || n >= (1ll << 31)) {
        bwt_set_error("bwt_sw_batch: bad argument");
        return BWT_ERR_ARG;
    }
    *cigar_total = 0;
    if (n == 0) return BWT_OK;
    BWT_TRY(index_select_device(device));
    int64_t qb = q_offsets[0], rb = r_offsets[0], qbytes = q_offsets[n] - qb, rbytes = r_offsets[n] - rb;
    std::vector<int64_t> qo((size_t)n + 1), ro((size_t)n + 1);
    int max_lq = 0;
    int64_t words = 0;
    for (int64_t k = 0; k <= n; k++) {
        qo[k] = q_offsets[k] - qb; ro[k] = r_offsets[k] - rb;
        if (k) {
            const int64_t l = qo[k] - qo[k - 1];
            if (l < 0 || ro[k] < ro[k - 1]) { bwt_set_error("bwt_sw_batch: offsets must be non-decreasing"); return BWT_ERR_ARG; }
            max_lq = (int)std::max<int64_t>(max_lq, std::min<int64_t>(l, 1 << 30));
            if (l <= BWT_PACK_MAX) words += (l + 15) >> 4;
        }
    }
    SwStage sw;
    sw.device = device;
    uint8_t *d_q = nullptr, *d_r = nullptr;
    int64_t *d_qo = nullptr, *d_ro = nullptr;
    uint32_t *d_off = nullptr, *d_tmp = nullptr, *d_pk = nullptr, *d_woff = nullptr, *d_len = nullptr, *d_exc = nullptr, *d_ref2 = nullptr, *d_refx = nullptr, *d_info = nullptr;
    int32_t *d_status = nullptr, *d_score = nullptr, *d_len_out = nullptr;
    int64_t *d_coff = nullptr;
    char *d_cig = nullptr;
    cudaStream_t stream = 0;
    int st = BWT_OK;
    do {
        if ((st = sw_stage_reserve(sw, (uint32_t)n)) != BWT_OK) break;
        if (cudaMalloc(&d_q, (size_t)qbytes + 16) || cudaMalloc(&d_r, (size_t)rbytes + 32) || cudaMalloc(&d_qo, (size_t)(n + 1) * 8) ||
            cudaMalloc(&d_ro, (size_t)(n + 1) * 8) || cudaMalloc(&d_off, (size_t)(n + 1) * 4) ||
            cudaMalloc(&d_tmp, scan_tmp_words((size_t)n + 1) * 4) || cudaMalloc(&d_status, (size_t)n * 4) ||
            cudaMalloc(&d_score, (size_t)n * 4) || cudaMalloc(&d_len_out, (size_t)n * 4) || cudaMalloc(&d_coff, (size_t)n * 8) ||
            cudaMalloc(&d_pk, (size_t)(words + 2) * 4) || cudaMalloc(&d_woff, (size_t)(n + 1) * 4) || cudaMalloc(&d_len, (size_t)(n + 1) * 4) ||
            cudaMalloc(&d_exc, (size_t)(n / 32 + 2) * 4) || cudaMalloc(&d_info, 16)) {
            bwt_set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_NOMEM; break;
        }
        if (qbytes) cudaMemcpy(d_q, q + qb, (size_t)qbytes, cudaMemcpyHostToDevice);
        cudaMemset(d_r, 'A', (size_t)rbytes + 32);
        if (rbytes) cudaMemcpy(d_r, r + rb, (size_t)rbytes, cudaMemcpyHostToDevice);
        cudaMemcpy(d_qo, qo.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_ro, ro.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
        const uint32_t n32 = (uint32_t)n;
        uint32_t info0[4] = {n32, 0, 0, 0};
        cudaMemcpy(d_info, info0, 16, cudaMemcpyHostToDevice);
        cudaMemset(d_pk, 0, (size_t)(words + 2) * 4);
        // the queries become the batch's "reads", the windows' buffer its window source
        if ((st = reads_geom_from_ascii(d_qo, n, d_len, d_woff, d_tmp, d_info + 2, stream, nullptr)) != BWT_OK) break;
        if ((st = reads_pack_device(d_q, d_qo, 0, n, d_len, d_woff, d_pk, d_exc, stream, nullptr)) != BWT_OK) break;
        if ((st = pack_ref_device(d_r, (uint64_t)rbytes, &d_ref2, &d_refx, stream)) != BWT_OK) break;
        DevReads R{};
        R.pk = d_pk; R.woff = d_woff; R.len = d_len; R.exc = d_exc; R.ascii = d_q; R.aoff = d_qo; R.exc_idx = nullptr;
        R.n_exc = 0; R.ascii_full = 1; R.aoff_base = 0; R.exc_base = 0;
        desc_from_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_qo, d_ro, n32, sw.d_desc, sw.d_out);
        { cudaError_t e0 = cudaGetLastError(); if (e0 == cudaSuccess && bwt_debug_sync()) e0 = cudaStreamSynchronize(stream);
          if (e0 != cudaSuccess) { bwt_set_error("desc_from_offsets_kernel: %s", cudaGetErrorString(e0)); st = BWT_ERR_CUDA; break; } }
        if ((st = sw_stage_run(sw, R, d_r, d_ref2, d_refx, d_info, n32, max_lq, stream)) != BWT_OK) break;
        cudaMemsetAsync(d_off + n, 0, 4, stream);
        sw_pairs_emit_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sw.d_out, n32, d_off);
        if ((st = exclusive_scan_u32(d_off, (size_t)n + 1, d_tmp, nullptr, stream, nullptr)) != BWT_OK) break;
        uint32_t tot = 0;
        cudaMemcpy(&tot, d_off + n, 4, cudaMemcpyDeviceToHost);
        *cigar_total = tot;
        if (cudaMalloc(&d_cig, (size_t)tot + 16)) { bwt_set_error("device allocation failed"); st = BWT_ERR_NOMEM; break; }
        sw_pairs_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sw.d_out, sw.d_cig_tmp, n32, d_off, d_status,
                                                                            d_score, d_coff, d_len_out, d_cig);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { bwt_set_error("sw emit failed: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; }
        cudaMemcpy(status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(score, d_score, (size_t)n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cigar_off, d_coff, (size_t)n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(cigar_len, d_len_out, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if ((int64_t)tot > cigar_cap || !cigars) {
            if (tot) { bwt_set_error("CIGAR capacity too small: need %u bytes", tot); st = BWT_ERR_NOSPACE; break; }
        } else if (tot) cudaMemcpy(cigars, d_cig, tot, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_q); cudaFree(d_r); cudaFree(d_qo); cudaFree(d_ro); cudaFree(d_off); cudaFree(d_tmp); cudaFree(d_status);
    cudaFree(d_score); cudaFree(d_len_out); cudaFree(d_coff); cudaFree(d_cig); cudaFree(d_pk); cudaFree(d_woff); cudaFree(d_len);
    cudaFree(d_exc); cudaFree(d_ref2); cudaFree(d_refx); cudaFree(d_info);
    sw_stage_free(sw);
    return st;
}
