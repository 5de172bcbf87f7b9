// The worker's inner loop (worker_bwt.erl:73-108) as a resident batch pipeline, and the C ABI on top of it.
//
//   upload (H2D)  ->  sga_kernel  ->  [sga_slow_kernel]  ->  windows  ->  sw_classify / sw_fast / [sw_general]
//                 ->  emit (scan + gather into the caller-facing layout)  ->  download (D2H)
//
// One bwt_batch owns its stream, device buffers and pinned staging, so several batches can be in flight on one
// (immutable) index from different host threads.
#include <algorithm>
#include <cstddef>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>

#include "common.cuh"
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

// offsets arrive as the caller has them: find the longest read and check that they do not decrease (the rebase to
// 0 is a second kernel, so that nobody reads an element a neighbour has already rewritten)
__global__ void prep_offsets_kernel(const int64_t *__restrict__ offs, int64_t n, uint32_t *__restrict__ info) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t len = k < n ? offs[k + 1] - offs[k] : 0;
    uint32_t l = len < 0 ? 0u : (len > (1 << 30) ? (1u << 30) : (uint32_t)len);
    if (len < 0) atomicExch(info + 1, 1u);
    for (int d = 16; d > 0; d >>= 1) l = max(l, __shfl_down_sync(0xFFFFFFFFu, l, d));
    if ((threadIdx.x & 31) == 0 && l) atomicMax(info, l);
}

__global__ void rebase_offsets_kernel(int64_t *__restrict__ offs, int64_t n, int64_t base) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) offs[k] -= base;
}

// pipelined aligner: results of a chunk move to their place in the caller's arrays
__global__ void rebase_result_kernel(int64_t *__restrict__ hit_off64, int64_t n, bwt_hit *__restrict__ hits, int64_t n_hits,
                                     int64_t hit_base, int64_t cig_base) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) hit_off64[k] += hit_base;
    if (k < n_hits) hits[k].cigar_off += cig_base;
}

// windows of worker_bwt.erl:84-90: Ref_len = Lq + D, Start = S - Ref_len + 200 into the padded reference
__global__ void windows_kernel(const Cluster *__restrict__ cl, uint32_t n, const int64_t *__restrict__ offs,
                               uint64_t ref_len_padded, SwDesc *__restrict__ desc, SwOut *__restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Cluster c = cl[k];
    int64_t q0 = offs[c.read];
    int32_t lq = (int32_t)(offs[c.read + 1] - q0);
    int64_t ref_len = (int64_t)lq + c.d;
    int64_t start = (int64_t)c.s - ref_len + BWT_REF_EXT;
    SwDesc d;
    d.q_off = q0; d.r_off = start; d.lq = lq; d.lr = (int32_t)ref_len;
    desc[k] = d;
    SwOut o;
    o.score = 0; o.cigar_len = 0; o.cigar_off = 0;
    // <<_:Start_pos/bytes, Ref_seq:Ref_len/bytes, _/binary>> = Ref  -> badmatch outside the binary
    o.status = (start < 0 || (uint64_t)(start + ref_len) > ref_len_padded) ? BWT_HIT_WINDOW_OOR : -1;
    out[k] = o;
}

// per read: number of hits and CIGAR bytes (inputs of the two scans)
__global__ void emit_count_kernel(const uint32_t *__restrict__ ncl, const uint32_t *__restrict__ cl_base,
                                  const SwOut *__restrict__ out, int64_t n, uint32_t *__restrict__ hit_cnt,
                                  uint32_t *__restrict__ cig_cnt) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n) { hit_cnt[n] = 0; cig_cnt[n] = 0; return; }   // slot n turns into the total under the exclusive scan
    uint32_t c = ncl[r], b = cl_base[r], bytes = 0;
    for (uint32_t k = 0; k < c; k++) bytes += (uint32_t)out[b + k].cigar_len;
    hit_cnt[r] = c;
    cig_cnt[r] = bytes;
}

// per read: hits in ASCENDING position (the reference prepends while folding sga's descending list,
// worker_bwt.erl:84-97), CIGAR bytes packed in the same order
__global__ void emit_kernel(const uint32_t *__restrict__ ncl, const uint32_t *__restrict__ cl_base,
                            const Cluster *__restrict__ cl, const SwDesc *__restrict__ desc,
                            const SwOut *__restrict__ out, const char *__restrict__ cig_tmp, int64_t n,
                            const uint32_t *__restrict__ hit_off, const uint32_t *__restrict__ cig_off, int64_t shift,
                            int64_t *__restrict__ hit_off64, bwt_hit *__restrict__ hits, char *__restrict__ cigars) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n) { hit_off64[n] = hit_off[n]; return; }
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
        h.status = o.status;
        h.cigar_len = o.cigar_len;
        h.cigar_off = co;
        hits[ho + k] = h;
        for (int t = 0; t < o.cigar_len; t++) cigars[co + t] = cig_tmp[o.cigar_off + t];
        co += (uint32_t)o.cigar_len;
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
    d.q_off = qo[k]; d.r_off = ro[k]; d.lq = (int32_t)(qo[k + 1] - qo[k]); d.lr = (int32_t)(ro[k + 1] - ro[k]);
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
    status[k] = o.status; score[k] = o.score; cig_off[k] = off[k]; cig_len[k] = o.cigar_len;
    for (int t = 0; t < o.cigar_len; t++) cigars[off[k] + t] = cig_tmp[o.cigar_off + t];
}

}  // namespace

struct bwt_batch {
    bwt_index *ix = nullptr;
    cudaStream_t stream = nullptr;
    int64_t max_reads = 0, max_bytes = 0;
    // inputs
    uint8_t *d_reads = nullptr;
    int64_t *d_offs = nullptr;
    uint8_t *h_stage = nullptr;           // pinned staging for pageable callers
    size_t h_stage_bytes = 0;
    int64_t n = 0, n_bytes = 0;
    int max_read_len = 0;
    // per read
    int32_t *d_read_status = nullptr;
    uint32_t *d_ncl = nullptr, *d_cl_base = nullptr, *d_hit_off = nullptr, *d_cig_off = nullptr, *d_scan_tmp = nullptr;
    int64_t *d_hit_off64 = nullptr;
    // clusters
    Cluster *d_clusters = nullptr;        // dense list (what windows / emit read)
    Cluster *d_clusters_sparse = nullptr; // the sub-list regions sga_kernel appends to (launch_sga_finish packs them)
    size_t cl_cap = 0, cl_sparse_cap = 0;
    SlowRead *d_slow = nullptr;
    uint32_t slow_cap = 0;
    uint64_t *d_slow_off = nullptr;
    size_t slow_off_cap = 0;
    uint32_t *d_slow_scratch = nullptr;
    size_t slow_scratch_cap = 0;
    SeedCounters *d_ctr = nullptr, *h_ctr = nullptr;
    uint32_t *d_totals = nullptr, *h_totals = nullptr;   // [0] hits, [1] cigar bytes
    SwStage sw;
    // outputs
    bwt_hit *d_hits = nullptr;
    size_t hits_cap = 0;
    char *d_cigars = nullptr;
    size_t cigars_cap = 0;
    int64_t n_hits = 0, cigar_bytes = 0;
    bool has_result = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bwt_stats stats{};
    std::mutex *stage_token = nullptr;    // set by the pipelined aligner: one chunk at a time inside a stage
    cudaStream_t dl_stream = nullptr;     // the aligner's D2H copies: the next chunk's upload does not wait for them
    cudaEvent_t ev_run_done = nullptr, ev_dl0 = nullptr, ev_dl1 = nullptr;
    bool dl_pending = false;
};

static int batch_alloc(bwt_batch *b) {
    int64_t nr = b->max_reads;
    CU_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&b->dl_stream, cudaStreamNonBlocking));
    for (auto &e : b->ev) CU_TRY(cudaEventCreate(&e));
    CU_TRY(cudaEventCreate(&b->ev_run_done));
    CU_TRY(cudaEventCreate(&b->ev_dl0));
    CU_TRY(cudaEventCreate(&b->ev_dl1));
    CU_TRY(cudaMalloc(&b->d_reads, (size_t)b->max_bytes + 16));
    CU_TRY(cudaMalloc(&b->d_offs, (size_t)(nr + 1) * sizeof(int64_t)));
    CU_TRY(cudaMalloc(&b->d_read_status, (size_t)nr * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&b->d_ncl, (size_t)nr * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_cl_base, (size_t)nr * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_hit_off, (size_t)(nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_cig_off, (size_t)(nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_hit_off64, (size_t)(nr + 1) * sizeof(int64_t)));
    CU_TRY(cudaMalloc(&b->d_scan_tmp, scan_tmp_words((size_t)nr + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&b->d_ctr, sizeof(SeedCounters)));
    CU_TRY(cudaMallocHost(&b->h_ctr, sizeof(SeedCounters)));
    CU_TRY(cudaMalloc(&b->d_totals, 2 * sizeof(uint32_t)));
    CU_TRY(cudaMallocHost(&b->h_totals, 2 * sizeof(uint32_t)));
    b->slow_cap = (uint32_t)std::min<int64_t>(nr, 1 << 20);
    CU_TRY(cudaMalloc(&b->d_slow, (size_t)b->slow_cap * sizeof(SlowRead)));
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
    cudaFree(b->d_reads); cudaFree(b->d_offs); cudaFree(b->d_read_status); cudaFree(b->d_ncl); cudaFree(b->d_cl_base);
    cudaFree(b->d_hit_off); cudaFree(b->d_cig_off); cudaFree(b->d_hit_off64); cudaFree(b->d_scan_tmp);
    cudaFree(b->d_clusters); cudaFree(b->d_clusters_sparse); cudaFree(b->d_slow); cudaFree(b->d_slow_off); cudaFree(b->d_slow_scratch);
    cudaFree(b->d_ctr); cudaFree(b->d_totals); cudaFree(b->d_hits); cudaFree(b->d_cigars);
    if (b->h_ctr) cudaFreeHost(b->h_ctr);
    if (b->h_totals) cudaFreeHost(b->h_totals);
    if (b->h_stage) cudaFreeHost(b->h_stage);
    sw_stage_free(b->sw);
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    if (b->ev_run_done) cudaEventDestroy(b->ev_run_done);
    if (b->ev_dl0) cudaEventDestroy(b->ev_dl0);
    if (b->ev_dl1) cudaEventDestroy(b->ev_dl1);
    if (b->dl_stream) { cudaStreamSynchronize(b->dl_stream); cudaStreamDestroy(b->dl_stream); }
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return BWT_OK;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int bwt_batch_upload(bwt_batch *b, const uint8_t *reads, const int64_t *offsets, int64_t n) {
    if (!b || !offsets || n < 0 || (n > 0 && !reads)) { bwt_set_error("bwt_batch_upload: bad argument"); return BWT_ERR_ARG; }
    if (n > b->max_reads) { bwt_set_error("batch of %lld reads exceeds max_reads %lld", (long long)n, (long long)b->max_reads); return BWT_ERR_ARG; }
    int64_t base = offsets[0], bytes = offsets[n] - base;
    if (bytes < 0 || bytes > b->max_bytes) { bwt_set_error("batch of %lld read bytes exceeds max_read_bytes %lld", (long long)bytes, (long long)b->max_bytes); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(b->ix->device));
    // pageable callers are staged through pinned memory; the offsets are rebased and checked on the device
    const bool pin_r = bytes == 0 || is_pinned(reads + base), pin_o = is_pinned(offsets);
    size_t need = (pin_r ? 0 : (size_t)bytes) + (pin_o ? 0 : (size_t)(n + 1) * sizeof(int64_t));
    if (need > b->h_stage_bytes) {
        if (b->h_stage) cudaFreeHost(b->h_stage);
        b->h_stage = nullptr;
        b->h_stage_bytes = need + need / 4 + 4096;
        CU_TRY(cudaMallocHost(&b->h_stage, b->h_stage_bytes));
    }
    const int64_t *src_o = offsets;
    if (!pin_o) { memcpy(b->h_stage, offsets, (size_t)(n + 1) * sizeof(int64_t)); src_o = reinterpret_cast<const int64_t *>(b->h_stage); }
    CU_TRY(cudaEventRecord(b->ev[4], b->stream));
    CU_TRY(cudaMemcpyAsync(b->d_offs, src_o, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, b->stream));
    CU_TRY(cudaMemsetAsync(b->d_totals, 0, 2 * sizeof(uint32_t), b->stream));
    unsigned blocks = (unsigned)((n + 1 + 255) / 256);
    prep_offsets_kernel<<<blocks, 256, 0, b->stream>>>(b->d_offs, n, b->d_totals);
    rebase_offsets_kernel<<<blocks, 256, 0, b->stream>>>(b->d_offs, n, base);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(b->h_totals, b->d_totals, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
    if (bytes > 0) {
        const uint8_t *src = reads + base;
        if (!pin_r) {
            uint8_t *sr = b->h_stage + (pin_o ? 0 : (size_t)(n + 1) * sizeof(int64_t));
            memcpy(sr, src, (size_t)bytes);
            src = sr;
        }
        CU_TRY(cudaMemcpyAsync(b->d_reads, src, (size_t)bytes, cudaMemcpyHostToDevice, b->stream));
    }
    CU_TRY(cudaEventRecord(b->ev[5], b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    if (b->h_totals[1]) { bwt_set_error("offsets must be non-decreasing"); return BWT_ERR_ARG; }
    int maxlen = (int)b->h_totals[0];
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev[4], b->ev[5]);
    b->stats.ms_h2d = ms;
    b->n = n; b->n_bytes = bytes; b->max_read_len = maxlen;
    b->has_result = false;
    return BWT_OK;
}

// seeding stage with capacity retries; on return h_ctr holds the counters and clusters are complete
static int run_seed_stage(bwt_batch *b) {
    DevIndex ix = b->ix->dv;
    // one lookup for the short subsequence at the front of the reads (8 + Lq mod 8 characters) as well
    const int tail = b->max_read_len >= 2 * BWT_MIN_LEN ? BWT_MIN_LEN + b->max_read_len % BWT_MIN_LEN : 0;
    if (const uint2 *t = index_tail_table(b->ix, tail)) { ix.kmer_tail = t; ix.tail_k = tail; }
    for (int attempt = 0; attempt < 6; attempt++) {
        BWT_TRY(dev_grow(b->d_clusters_sparse, b->cl_sparse_cap, b->cl_cap, false));
        const uint32_t cl_cap = (uint32_t)std::min<size_t>(std::min(b->cl_cap, b->cl_sparse_cap), 0xFFFFFFF0u);
        CU_TRY(cudaMemsetAsync(b->d_ctr, 0, sizeof(SeedCounters), b->stream));
        CU_TRY(cudaEventRecord(b->ev[6], b->stream));
        BWT_TRY(launch_sga(ix, b->d_reads, b->d_offs, b->n, b->max_read_len, b->d_read_status, b->d_ncl, b->d_cl_base,
                           b->d_clusters_sparse, cl_cap, b->d_slow, b->slow_cap, b->d_ctr, b->stream));
        CU_TRY(cudaEventRecord(b->ev[7], b->stream));
        b->stats.launches += 1;
        // only the slow-read count is needed on the host now (the first 64 bytes of the counters)
        CU_TRY(cudaMemcpyAsync(b->h_ctr, b->d_ctr, offsetof(SeedCounters, dense_base), cudaMemcpyDeviceToHost, b->stream));
        CU_TRY(cudaStreamSynchronize(b->stream));
        cudaEventElapsedTime(&b->stats.ms_sga_kernel, b->ev[6], b->ev[7]);
        uint32_t n_slow = b->h_ctr->n_slow;
        if (n_slow > b->slow_cap) {          // cannot happen with slow_cap = max_reads unless capped at 2^20
            cudaFree(b->d_slow); b->d_slow = nullptr;
            b->slow_cap = n_slow + 1024;
            CU_TRY(cudaMalloc(&b->d_slow, (size_t)b->slow_cap * sizeof(SlowRead)));
            continue;
        }
        if (n_slow > 0 && !b->h_ctr->cluster_overflow) {
            std::vector<SlowRead> sl(n_slow);
            CU_TRY(cudaMemcpy(sl.data(), b->d_slow, n_slow * sizeof(SlowRead), cudaMemcpyDeviceToHost));
            std::vector<uint64_t> off(n_slow + 1);
            off[0] = 0;
            for (uint32_t k = 0; k < n_slow; k++) {
                uint64_t p = 1;
                while (p < sl[k].seeds) p <<= 1;
                off[k + 1] = off[k] + p;
            }
            BWT_TRY(dev_grow(b->d_slow_off, b->slow_off_cap, (size_t)n_slow + 1));
            BWT_TRY(dev_grow(b->d_slow_scratch, b->slow_scratch_cap, (size_t)off[n_slow]));
            CU_TRY(cudaMemcpyAsync(b->d_slow_off, off.data(), (n_slow + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
            BWT_TRY(launch_sga_slow(ix, b->d_reads, b->d_offs, b->d_slow, b->d_slow_off, n_slow, b->d_slow_scratch,
                                    b->d_read_status, b->d_ncl, b->d_cl_base, b->d_clusters_sparse, cl_cap, b->d_ctr, b->stream));
            b->stats.launches += 1;
        }
        // sub-lists -> dense cluster list, dense cl_base, n_clusters
        BWT_TRY(launch_sga_finish(b->d_clusters_sparse, b->d_clusters, cl_cap, b->d_cl_base, b->d_ncl, b->n, b->d_ctr, b->stream,
                                  &b->stats.launches));
        CU_TRY(cudaMemcpyAsync(b->h_ctr, b->d_ctr, offsetof(SeedCounters, dense_base), cudaMemcpyDeviceToHost, b->stream));
        CU_TRY(cudaStreamSynchronize(b->stream));
        if (!b->h_ctr->cluster_overflow) return BWT_OK;
        // a sub-list ran out of room: n_clusters is a lower bound of what is needed (refused reads not counted in full);
        // the regions are equal shares, so ask for twice the total
        size_t want = (size_t)std::max<uint64_t>(b->h_ctr->n_clusters, b->cl_cap) * 2 + 4096;
        {   // and every region must hold the fullest sub-list (one read with very many clusters fills a single region)
            std::vector<uint32_t> sub((size_t)SGA_SUBLISTS * 32);
            CU_TRY(cudaMemcpy(sub.data(), reinterpret_cast<const uint8_t *>(b->d_ctr) + offsetof(SeedCounters, sub_n),
                              sub.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            uint64_t fullest = 0;
            for (int r = 0; r < SGA_SUBLISTS; r++) fullest = std::max<uint64_t>(fullest, sub[(size_t)32 * r]);
            want = std::max<size_t>(want, (size_t)(fullest + fullest / 4 + 64) * SGA_SUBLISTS);
        }
        if (want * sizeof(Cluster) > (48ull << 30)) { bwt_set_error("cluster list beyond 48 GiB: use smaller batches"); return BWT_ERR_NOMEM; }
        BWT_TRY(dev_grow(b->d_clusters, b->cl_cap, want, false));
    }
    bwt_set_error("cluster list kept overflowing");
    return BWT_ERR_NOMEM;
}

static int run_emit_stage(bwt_batch *b, uint32_t ncl) {
    int64_t n = b->n;
    unsigned blocks = (unsigned)((n + 1 + 255) / 256);
    emit_count_kernel<<<blocks, 256, 0, b->stream>>>(b->d_ncl, b->d_cl_base, b->sw.d_out, n, b->d_hit_off, b->d_cig_off);
    b->stats.launches += 1;
    CU_TRY(cudaGetLastError());
    BWT_TRY(exclusive_scan_u32(b->d_hit_off, (size_t)n + 1, b->d_scan_tmp, nullptr, b->stream, &b->stats.launches));
    BWT_TRY(exclusive_scan_u32(b->d_cig_off, (size_t)n + 1, b->d_scan_tmp, nullptr, b->stream, &b->stats.launches));
    // element n of an exclusive scan over n+1 items is the total (item n is garbage-free: written as 0 below)
    BWT_TRY(dev_grow(b->d_hits, b->hits_cap, (size_t)ncl));
    BWT_TRY(dev_grow(b->d_cigars, b->cigars_cap, (size_t)b->sw.h_ctr->cig_used + 16));
    emit_kernel<<<blocks, 256, 0, b->stream>>>(b->d_ncl, b->d_cl_base, b->d_clusters, b->sw.d_desc, b->sw.d_out,
                                               b->sw.d_cig_tmp, n, b->d_hit_off, b->d_cig_off, b->ix->shift,
                                               b->d_hit_off64, b->d_hits, b->d_cigars);
    b->stats.launches += 1;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(b->h_totals, b->d_hit_off + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaMemcpyAsync(b->h_totals + 1, b->d_cig_off + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
    return BWT_OK;
}

static int run_core(bwt_batch *b, bool with_sw) {
    if (!b) { bwt_set_error("null batch"); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(b->ix->device));
    float h2d = b->stats.ms_h2d;
    memset(&b->stats, 0, sizeof b->stats);
    b->stats.ms_h2d = h2d;
    b->sw.launches = 0;
    b->n_hits = 0; b->cigar_bytes = 0;
    // The aligner's chunks take the token stage by stage: while one chunk's host thread sits between two stages (reading
    // counters, sizing the next launch), another chunk's stage runs instead of the GPU idling.
    struct Token {
        std::mutex *m;
        explicit Token(std::mutex *mu) : m(mu) { if (m) m->lock(); }
        ~Token() { if (m) m->unlock(); }
    };
    CU_TRY(cudaEventRecord(b->ev[0], b->stream));
    if (b->n > 0) { Token tk(b->stage_token); BWT_TRY(run_seed_stage(b)); }
    else memset(b->h_ctr, 0, sizeof(SeedCounters));
    CU_TRY(cudaEventRecord(b->ev[1], b->stream));
    uint32_t ncl = b->n > 0 ? b->h_ctr->n_clusters : 0;
    if (with_sw) {
        BWT_TRY(sw_stage_reserve(b->sw, (uint32_t)b->cl_cap));
        if (ncl > 0) {
            Token tk(b->stage_token);
            windows_kernel<<<(ncl + 255) / 256, 256, 0, b->stream>>>(b->d_clusters, ncl, b->d_offs, b->ix->dv.ref_len_padded,
                                                                     b->sw.d_desc, b->sw.d_out);
            b->stats.launches += 1;
            K_CHECK(b->stream, "windows_kernel");
            BWT_TRY(sw_stage_run(b->sw, b->d_reads, b->ix->dv.ref, ncl, b->max_read_len, b->stream));
        } else {
            memset(b->sw.h_ctr, 0, sizeof(SwCounters));
        }
        CU_TRY(cudaEventRecord(b->ev[2], b->stream));
        {
            Token tk(b->stage_token);
            if (b->n > 0) BWT_TRY(run_emit_stage(b, ncl));
            CU_TRY(cudaEventRecord(b->ev[3], b->stream));
            CU_TRY(cudaStreamSynchronize(b->stream));
        }
        b->n_hits = b->n > 0 ? b->h_totals[0] : 0;
        b->cigar_bytes = b->n > 0 ? b->h_totals[1] : 0;
    } else {
        CU_TRY(cudaEventRecord(b->ev[2], b->stream));
        CU_TRY(cudaEventRecord(b->ev[3], b->stream));
        CU_TRY(cudaStreamSynchronize(b->stream));
        b->n_hits = ncl;
    }
    bwt_stats &s = b->stats;
    s.reads = (uint64_t)b->n;
    s.searches = b->h_ctr->searches; s.lf_steps = b->h_ctr->lf_steps; s.hbm_lf_steps = b->h_ctr->hbm_lf_steps;
    s.located_rows = b->h_ctr->located;
    s.ms_sw_kernel = with_sw ? b->sw.ms_fast_kernel : 0.f;
    s.clusters = ncl;
    s.sw_calls = b->sw.sw_calls; s.sw_cells = b->sw.cells; s.sw_general = b->sw.n_general;
    s.sw_shortcut = b->sw.shortcut_calls; s.sw_shortcut_cells = b->sw.shortcut_cells;
    s.launches += b->sw.launches;
    cudaEventElapsedTime(&s.ms_seed, b->ev[0], b->ev[1]);
    cudaEventElapsedTime(&s.ms_sw, b->ev[1], b->ev[2]);
    cudaEventElapsedTime(&s.ms_emit, b->ev[2], b->ev[3]);
    cudaEventElapsedTime(&s.ms_total, b->ev[0], b->ev[3]);
    b->has_result = with_sw;
    return BWT_OK;
}

extern "C" int bwt_batch_run(bwt_batch *b) { return run_core(b, true); }

extern "C" int bwt_batch_result_size(bwt_batch *b, int64_t *n_hits, int64_t *cigar_bytes) {
    if (!b) { bwt_set_error("null batch"); return BWT_ERR_ARG; }
    if (n_hits) *n_hits = b->n_hits;
    if (cigar_bytes) *cigar_bytes = b->cigar_bytes;
    return BWT_OK;
}

extern "C" int bwt_batch_download(bwt_batch *b, int32_t *read_status, int64_t *hit_offsets, bwt_hit *hits,
                                  int64_t hits_cap, char *cigars, int64_t cigar_cap) {
    if (!b || !b->has_result) { bwt_set_error("bwt_batch_download: no result (call bwt_batch_run first)"); return BWT_ERR_ARG; }
    if (hits_cap < b->n_hits || cigar_cap < b->cigar_bytes) {
        bwt_set_error("output capacity too small: need %lld hits and %lld CIGAR bytes", (long long)b->n_hits, (long long)b->cigar_bytes);
        return BWT_ERR_NOSPACE;
    }
    CU_TRY(cudaSetDevice(b->ix->device));
    CU_TRY(cudaEventRecord(b->ev[4], b->stream));
    if (read_status && b->n) CU_TRY(cudaMemcpyAsync(read_status, b->d_read_status, (size_t)b->n * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    if (hit_offsets) {
        if (b->n) CU_TRY(cudaMemcpyAsync(hit_offsets, b->d_hit_off64, (size_t)(b->n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, b->stream));
        else hit_offsets[0] = 0;
    }
    if (hits && b->n_hits) CU_TRY(cudaMemcpyAsync(hits, b->d_hits, (size_t)b->n_hits * sizeof(bwt_hit), cudaMemcpyDeviceToHost, b->stream));
    if (cigars && b->cigar_bytes) CU_TRY(cudaMemcpyAsync(cigars, b->d_cigars, (size_t)b->cigar_bytes, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaEventRecord(b->ev[5], b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    cudaEventElapsedTime(&b->stats.ms_d2h, b->ev[4], b->ev[5]);
    return BWT_OK;
}

extern "C" int bwt_batch_stats(bwt_batch *b, bwt_stats *out) {
    if (!b || !out) { bwt_set_error("bwt_batch_stats: null argument"); return BWT_ERR_ARG; }
    *out = b->stats;
    return BWT_OK;
}

extern "C" int bwt_batch_device_result(bwt_batch *b, void **d_hit_offsets, void **d_hits, void **d_cigars) {
    if (!b || !b->has_result) { bwt_set_error("bwt_batch_device_result: no result"); return BWT_ERR_ARG; }
    if (d_hit_offsets) *d_hit_offsets = b->d_hit_off64;
    if (d_hits) *d_hits = b->d_hits;
    if (d_cigars) *d_cigars = b->d_cigars;
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
// that batch's own stream, so the H2D of one chunk, the kernels of another and the D2H of a third overlap.  Results
// land in the caller's arrays at their final place: a chunk learns its bases once every earlier chunk has finished.
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
    std::condition_variable cv;
    int64_t next_chunk = 0;          // chunks [0, next_chunk) have published their totals
    int64_t hit_base = 0, cig_base = 0;
    bool space = true;
    int status = BWT_OK;
    std::string error;
};

static void stats_add(bwt_stats &t, const bwt_stats &s) {
    t.reads += s.reads; t.searches += s.searches; t.lf_steps += s.lf_steps; t.hbm_lf_steps += s.hbm_lf_steps;
    t.located_rows += s.located_rows; t.clusters += s.clusters; t.sw_calls += s.sw_calls; t.sw_cells += s.sw_cells;
    t.sw_shortcut += s.sw_shortcut; t.sw_shortcut_cells += s.sw_shortcut_cells;
    t.sw_general += s.sw_general; t.ms_total += s.ms_total; t.ms_seed += s.ms_seed; t.ms_sw += s.ms_sw; t.ms_emit += s.ms_emit;
    t.ms_h2d += s.ms_h2d; t.ms_d2h += s.ms_d2h; t.ms_sga_kernel += s.ms_sga_kernel; t.ms_sw_kernel += s.ms_sw_kernel;
    t.launches += s.launches;
}
}  // namespace

extern "C" int bwt_aligner_run(bwt_aligner *a, const uint8_t *reads, const int64_t *offsets, int64_t n, int32_t *read_status,
                               int64_t *hit_offsets, bwt_hit *hits, int64_t hits_cap, int64_t *n_hits, char *cigars,
                               int64_t cigar_cap, int64_t *cigar_bytes) {
    if (!a || !offsets || n < 0 || !hit_offsets || !n_hits || !cigar_bytes || (n > 0 && !reads)) { bwt_set_error("bwt_aligner_run: null argument"); return BWT_ERR_ARG; }
    memset(&a->stats, 0, sizeof a->stats);
    hit_offsets[0] = 0;
    *n_hits = 0; *cigar_bytes = 0;
    if (n == 0) return BWT_OK;
    std::vector<int64_t> bounds(1, 0);                      // chunk c = reads [bounds[c], bounds[c+1])
    for (int64_t lo = 0; lo < n;) {
        int64_t hi = chunk_end_lim(offsets, lo, n, a->chunk_reads, a->chunk_bytes);
        if (offsets[hi] - offsets[lo] > a->chunk_bytes) { bwt_set_error("a single read exceeds %lld bytes", (long long)a->chunk_bytes); return BWT_ERR_LIMIT; }
        bounds.push_back(hi);
        lo = hi;
    }
    const int64_t nchunks = (int64_t)bounds.size() - 1;
    const int depth = (int)std::min<int64_t>((int64_t)a->batches.size(), nchunks);
    PipeShared sh;
    std::vector<bwt_stats> wstats((size_t)depth);
    const bool trace = getenv("BWT_TRACE") != nullptr;       // per-chunk host timeline on stderr (debug aid)
    const char *env_tok = getenv("BWT_COMPUTE_TOKEN");
    const int token_mode = env_tok ? atoi(env_tok) : 2;
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
            if (!skip) {
                st = bwt_batch_upload(b, reads, offsets + lo, m);
                t1 = now_ms();
                if (st == BWT_OK && b->dl_pending) {     // the previous chunk's results must have left before this run overwrites them
                    cudaError_t e = cudaStreamSynchronize(b->dl_stream);
                    b->dl_pending = false;
                    if (e != cudaSuccess) { bwt_set_error("aligner download: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; }
                    else { float ms = 0; cudaEventElapsedTime(&ms, b->ev_dl0, b->ev_dl1); wstats[(size_t)w].ms_d2h += ms; }
                }
                if (st == BWT_OK) {
                    // BWT_COMPUTE_TOKEN: 2 (default) = one chunk at a time inside a whole run; 1 = inside a stage; 0 = free
                    // (measured on 10 M x 100 bp, depth 2 / 3: 289 / 283 M reads/s, 290 / 247, 256 / -)
                    std::unique_lock<std::mutex> run_token(sh.compute_mu, std::defer_lock);
                    if (token_mode == 2) run_token.lock();
                    b->stage_token = token_mode == 1 ? &sh.compute_mu : nullptr;
                    st = bwt_batch_run(b);
                    b->stage_token = nullptr;
                }
                t2 = now_ms();
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
                if (!(hits && cigars && sh.hit_base <= hits_cap && sh.cig_base <= cigar_cap)) sh.space = false;
                room = sh.space && sh.status == BWT_OK;
                sh.next_chunk = c + 1;
            }
            sh.cv.notify_all();
            t3 = now_ms();
            if (skip || st != BWT_OK) continue;
            if (room) {
                const int64_t cnt = std::max(m + 1, nh);
                rebase_result_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, b->stream>>>(b->d_hit_off64, m, b->d_hits, nh, hb, cb);
                // entry m of this chunk equals entry 0 of the next one: only the last chunk writes it
                st = cudaGetLastError() == cudaSuccess ? BWT_OK : BWT_ERR_CUDA;
                if (st == BWT_OK) {
                    // D2H on the batch's download stream, behind the rebase kernel; not waited for here: the worker goes on
                    // to upload its next chunk (other buffers) and collects the copy before its next run
                    cudaEventRecord(b->ev_run_done, b->stream);
                    cudaStream_t s = b->dl_stream;
                    cudaStreamWaitEvent(s, b->ev_run_done, 0);
                    cudaEventRecord(b->ev_dl0, s);
                    if (read_status) cudaMemcpyAsync(read_status + lo, b->d_read_status, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
                    cudaMemcpyAsync(hit_offsets + lo, b->d_hit_off64, (size_t)(m + (c + 1 == nchunks ? 1 : 0)) * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
                    if (nh) cudaMemcpyAsync(hits + hb, b->d_hits, (size_t)nh * sizeof(bwt_hit), cudaMemcpyDeviceToHost, s);
                    if (nc) cudaMemcpyAsync(cigars + cb, b->d_cigars, (size_t)nc, cudaMemcpyDeviceToHost, s);
                    cudaEventRecord(b->ev_dl1, s);
                    if (cudaGetLastError() != cudaSuccess) { bwt_set_error("aligner download: enqueue failed"); st = BWT_ERR_CUDA; }
                    else b->dl_pending = true;
                }
                if (st != BWT_OK) { std::lock_guard<std::mutex> g(sh.mu); if (sh.status == BWT_OK) { sh.status = st; sh.error = bwt_last_error(); } }
            }
            stats_add(wstats[(size_t)w], b->stats);
            if (trace)
                fprintf(stderr, "chunk %3lld w%d: upload %.2f-%.2f run -%.2f (seed %.2f sw %.2f emit %.2f) wait -%.2f download -%.2f\n", (long long)c, w,
                        t0, t1, t2, b->stats.ms_seed, b->stats.ms_sw, b->stats.ms_emit, t3, now_ms());
        }
    };
    auto worker_all = [&](int w) {
        worker(w);
        bwt_batch *b = a->batches[(size_t)w];
        if (b->dl_pending) {
            cudaError_t e = cudaStreamSynchronize(b->dl_stream);
            b->dl_pending = false;
            if (e != cudaSuccess) { std::lock_guard<std::mutex> g(sh.mu); if (sh.status == BWT_OK) { sh.status = BWT_ERR_CUDA; sh.error = cudaGetErrorString(e); } }
            else { float ms = 0; cudaEventElapsedTime(&ms, b->ev_dl0, b->ev_dl1); wstats[(size_t)w].ms_d2h += ms; }
        }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < depth; w++) th.emplace_back(worker_all, w);
    worker_all(0);
    for (auto &t : th) t.join();
    for (int w = 0; w < depth; w++) stats_add(a->stats, wstats[(size_t)w]);
    *n_hits = sh.hit_base; *cigar_bytes = sh.cig_base;
    if (sh.status != BWT_OK) { bwt_set_error("%s", sh.error.c_str()); return sh.status; }
    if (!sh.space) {
        bwt_set_error("output capacity too small: need %lld hits and %lld CIGAR bytes", (long long)sh.hit_base, (long long)sh.cig_base);
        return BWT_ERR_NOSPACE;
    }
    return BWT_OK;
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
    const int depth = n > 2 * CHUNK_READS ? 3 : (n > CHUNK_READS ? 2 : 1);
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
        if ((st = run_core(b, false)) != BWT_OK) break;
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
    CU_TRY(cudaMalloc(&d_subs, (size_t)n * sub_len));
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
    for (int64_t k = 0; k <= n; k++) {
        qo[k] = q_offsets[k] - qb; ro[k] = r_offsets[k] - rb;
        if (k) max_lq = (int)std::max<int64_t>(max_lq, std::min<int64_t>(qo[k] - qo[k - 1], 1 << 30));
    }
    SwStage sw;
    sw.device = device;
    uint8_t *d_q = nullptr, *d_r = nullptr;
    int64_t *d_qo = nullptr, *d_ro = nullptr;
    uint32_t *d_off = nullptr, *d_tmp = nullptr;
    int32_t *d_status = nullptr, *d_score = nullptr, *d_len = nullptr;
    int64_t *d_coff = nullptr;
    char *d_cig = nullptr;
    cudaStream_t stream = 0;
    int st = BWT_OK;
    do {
        if ((st = sw_stage_reserve(sw, (uint32_t)n)) != BWT_OK) break;
        if (cudaMalloc(&d_q, (size_t)qbytes + 16) || cudaMalloc(&d_r, (size_t)rbytes + 16) || cudaMalloc(&d_qo, (size_t)(n + 1) * 8) ||
            cudaMalloc(&d_ro, (size_t)(n + 1) * 8) || cudaMalloc(&d_off, (size_t)(n + 1) * 4) ||
            cudaMalloc(&d_tmp, scan_tmp_words((size_t)n + 1) * 4) || cudaMalloc(&d_status, (size_t)n * 4) ||
            cudaMalloc(&d_score, (size_t)n * 4) || cudaMalloc(&d_len, (size_t)n * 4) || cudaMalloc(&d_coff, (size_t)n * 8)) {
            bwt_set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_NOMEM; break;
        }
        if (qbytes) cudaMemcpy(d_q, q + qb, (size_t)qbytes, cudaMemcpyHostToDevice);
        if (rbytes) cudaMemcpy(d_r, r + rb, (size_t)rbytes, cudaMemcpyHostToDevice);
        cudaMemcpy(d_qo, qo.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_ro, ro.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
        desc_from_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_qo, d_ro, (uint32_t)n, sw.d_desc, sw.d_out);
        { cudaError_t e0 = cudaGetLastError(); if (e0 == cudaSuccess && bwt_debug_sync()) e0 = cudaStreamSynchronize(stream);
          if (e0 != cudaSuccess) { bwt_set_error("desc_from_offsets_kernel: %s", cudaGetErrorString(e0)); st = BWT_ERR_CUDA; break; } }
        if ((st = sw_stage_run(sw, d_q, d_r, (uint32_t)n, max_lq, stream)) != BWT_OK) break;
        cudaMemsetAsync(d_off + n, 0, 4, stream);
        sw_pairs_emit_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sw.d_out, (uint32_t)n, d_off);
        if ((st = exclusive_scan_u32(d_off, (size_t)n + 1, d_tmp, nullptr, stream, nullptr)) != BWT_OK) break;
        uint32_t tot = 0;
        cudaMemcpy(&tot, d_off + n, 4, cudaMemcpyDeviceToHost);
        *cigar_total = tot;
        if (cudaMalloc(&d_cig, (size_t)tot + 16)) { bwt_set_error("device allocation failed"); st = BWT_ERR_NOMEM; break; }
        sw_pairs_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sw.d_out, sw.d_cig_tmp, (uint32_t)n, d_off, d_status,
                                                                            d_score, d_coff, d_len, d_cig);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { bwt_set_error("sw emit failed: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; }
        cudaMemcpy(status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(score, d_score, (size_t)n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cigar_off, d_coff, (size_t)n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(cigar_len, d_len, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if ((int64_t)tot > cigar_cap || !cigars) {
            if (tot) { bwt_set_error("CIGAR capacity too small: need %u bytes", tot); st = BWT_ERR_NOSPACE; break; }
        } else if (tot) cudaMemcpy(cigars, d_cig, tot, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d_q); cudaFree(d_r); cudaFree(d_qo); cudaFree(d_ro); cudaFree(d_off); cudaFree(d_tmp); cudaFree(d_status);
    cudaFree(d_score); cudaFree(d_len); cudaFree(d_coff); cudaFree(d_cig);
    sw_stage_free(sw);
    return st;
}
