// Reads on the device: the 2-bit packed, reversed layout every kernel of the path works on (DevReads, common.cuh),
// produced either on the device from an ASCII batch (bwt_batch_upload) or on the host (bwt_pack_reads: what a FASTQ
// ingest hands to bwt_aligner_run_packed, 4x fewer bytes over PCIe).  A read with any byte outside ACGT, or longer
// than BWT_PACK_MAX, is flagged and keeps its ASCII text: sga:skip_Ns/1, the 1.6 score of 'N' (bwt.hrl:15) and the
// case_clause of bwa.erl:15-24 are decided on the bytes themselves.
#include <string.h>
#include <thread>
#include <vector>

#include "common.cuh"
#include "reads.cuh"

namespace {

__device__ __forceinline__ uint32_t words_of(uint32_t len) { return len > BWT_PACK_MAX ? 0u : (len + 15u) >> 4; }

// lengths and word counts from ASCII offsets; info[0] = longest read (atomicMax), info[1] = 1 when offsets decrease
__global__ void ascii_geom_kernel(const int64_t *__restrict__ aoff, int64_t n, uint32_t *__restrict__ len,
                                  uint32_t *__restrict__ nw, uint32_t *__restrict__ info) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n) { nw[n] = 0; return; }
    const int64_t l = aoff[r + 1] - aoff[r];
    if (l < 0) atomicExch(info + 1, 1u);
    const uint32_t lq = l < 0 ? 0u : (l > (1 << 30) ? (1u << 30) : (uint32_t)l);
    len[r] = lq;
    nw[r] = words_of(lq);
}

__global__ void fixed_geom_kernel(int64_t n, uint32_t fixed_len, uint32_t *__restrict__ len, uint32_t *__restrict__ woff) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    const uint32_t w = words_of(fixed_len);
    woff[r] = (uint32_t)r * w;
    if (r < n) len[r] = fixed_len;
}

__global__ void len16_geom_kernel(const uint16_t *__restrict__ l16, int64_t n, uint32_t *__restrict__ len, uint32_t *__restrict__ nw) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    if (r == n) { nw[n] = 0; return; }
    const uint32_t lq = l16[r];
    len[r] = lq;
    nw[r] = words_of(lq);
}

// one thread per read: ASCII -> reversed 2-bit words; the warp's ballot is the read's word of the exception mask
__global__ void pack_reads_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ aoff, int64_t aoff_base, int64_t n,
                                  const uint32_t *__restrict__ len, const uint32_t *__restrict__ woff,
                                  uint32_t *__restrict__ pk, uint32_t *__restrict__ exc) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool flagged = false;
    if (r < n) {
        const int lq = (int)len[r];
        if (lq > BWT_PACK_MAX) flagged = true;
        else {
            const uint8_t *q = ascii + (aoff[r] - aoff_base);
            uint32_t *o = pk + woff[r];
            const int nw = (lq + 15) >> 4;
            bool valid = true;
            for (int w = 0; w < nw; w++) {
                // word w holds the 16 bases ending at e = lq-1-16w: packed forward four bytes at a time, then turned round
                const int e = lq - 1 - 16 * w, start = max(e - 15, 0), cnt = e - start + 1;
                uint32_t fw = pack16_acgt(q + start, cnt, valid);        // base start+t at bits 2t, codes A0 C1 T2 G3
                fw ^= (fw >> 1) & 0x55555555u;                           // -> A0 C1 G2 T3
                fw <<= 2 * (16 - cnt);                                   // a short first word: bases to the top
                o[w] = rev2(fw);
            }
            flagged = !valid;
        }
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, flagged);
    if ((threadIdx.x & 31) == 0 && r < ((n + 31) & ~(int64_t)31)) exc[r >> 5] = m;
}

__global__ void exc_scatter_kernel(const uint32_t *__restrict__ idx, uint32_t n_exc, uint32_t base, uint32_t *__restrict__ exc) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_exc) { const uint32_t r = idx[k] - base; atomicOr(exc + (r >> 5), 1u << (r & 31)); }
}

}  // namespace

int reads_geom_from_ascii(const int64_t *d_aoff, int64_t n, uint32_t *d_len, uint32_t *d_woff, uint32_t *d_scan_tmp,
                          uint32_t *d_info, cudaStream_t stream, uint32_t *launches) {
    ascii_geom_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, stream>>>(d_aoff, n, d_len, d_woff, d_info);
    K_CHECK(stream, "ascii_geom_kernel");
    if (launches) *launches += 1;
    return exclusive_scan_u32(d_woff, (size_t)n + 1, d_scan_tmp, nullptr, stream, launches);
}

int reads_geom_fixed(int64_t n, uint32_t fixed_len, uint32_t *d_len, uint32_t *d_woff, cudaStream_t stream, uint32_t *launches) {
    fixed_geom_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, stream>>>(n, fixed_len, d_len, d_woff);
    CU_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return BWT_OK;
}

int reads_geom_from_len16(const uint16_t *d_len16, int64_t n, uint32_t *d_len, uint32_t *d_woff, uint32_t *d_scan_tmp,
                          cudaStream_t stream, uint32_t *launches) {
    len16_geom_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, stream>>>(d_len16, n, d_len, d_woff);
    CU_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return exclusive_scan_u32(d_woff, (size_t)n + 1, d_scan_tmp, nullptr, stream, launches);
}

int reads_pack_device(const uint8_t *d_ascii, const int64_t *d_aoff, int64_t aoff_base, int64_t n, const uint32_t *d_len, const uint32_t *d_woff,
                      uint32_t *d_pk, uint32_t *d_exc, cudaStream_t stream, uint32_t *launches) {
    if (n <= 0) return BWT_OK;
    const int64_t threads = (n + 31) & ~(int64_t)31;
    pack_reads_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(d_ascii, d_aoff, aoff_base, n, d_len, d_woff, d_pk, d_exc);
    K_CHECK(stream, "pack_reads_kernel");
    if (launches) *launches += 1;
    return BWT_OK;
}

int reads_exc_scatter(const uint32_t *d_exc_idx, uint32_t n_exc, uint32_t base, int64_t n, uint32_t *d_exc, cudaStream_t stream, uint32_t *launches) {
    CU_TRY(cudaMemsetAsync(d_exc, 0, (size_t)((n + 31) / 32 + 1) * sizeof(uint32_t), stream));
    if (n_exc) {
        exc_scatter_kernel<<<(n_exc + 255) / 256, 256, 0, stream>>>(d_exc_idx, n_exc, base, d_exc);
        CU_TRY(cudaGetLastError());
        if (launches) *launches += 1;
    }
    return BWT_OK;
}

// ---- host side: the same layout from ASCII reads ------------------------------------------------------------------

namespace {
inline int code_of(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// packs read (q, lq) into nw reversed words; false when a byte is outside ACGT
inline bool pack_one(const uint8_t *q, int lq, uint32_t *o) {
    const int nw = (lq + 15) >> 4;
    bool ok = true;
    for (int w = 0; w < nw; w++) {
        uint32_t x = 0;
        const int jhi = std::min(lq, 16 * w + 16);
        for (int j = 16 * w; j < jhi; j++) {
            int c = code_of(q[lq - 1 - j]);
            if (c < 0) { ok = false; c = 0; }
            x |= (uint32_t)c << (2 * (j & 15));
        }
        o[w] = x;
    }
    return ok;
}
}  // namespace

extern "C" int64_t bwt_pack_words(const int64_t *offsets, int64_t n) {
    if (!offsets || n < 0) return -1;
    int64_t w = 0;
    for (int64_t r = 0; r < n; r++) {
        const int64_t l = offsets[r + 1] - offsets[r];
        if (l < 0) return -1;
        if (l <= BWT_PACK_MAX) w += (l + 15) >> 4;
    }
    return w;
}

extern "C" int bwt_pack_reads(const uint8_t *reads, const int64_t *offsets, int64_t n, int threads, uint32_t *words,
                              int64_t words_cap, uint16_t *lens, uint32_t *exc_idx, int64_t exc_cap, int64_t *exc_off,
                              uint8_t *exc_ascii, int64_t exc_ascii_cap, int64_t *n_words, int64_t *n_exc,
                              int64_t *exc_bytes) {
    if (!offsets || n < 0 || (n > 0 && !reads) || !words || !lens || !n_words || !n_exc || !exc_bytes || n >= (1ll << 32)) {
        bwt_set_error("bwt_pack_reads: bad argument");
        return BWT_ERR_ARG;
    }
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n < 4096) threads = 1;
    // pass 1: word offsets of the thread ranges (contiguous read ranges)
    std::vector<int64_t> wbase((size_t)threads + 1, 0);
    for (int t = 0; t < threads; t++) {
        const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
        const int64_t w = bwt_pack_words(offsets + lo, hi - lo);
        if (w < 0) { bwt_set_error("bwt_pack_reads: offsets must be non-decreasing"); return BWT_ERR_ARG; }
        wbase[(size_t)t + 1] = wbase[(size_t)t] + w;
    }
    *n_words = wbase[(size_t)threads];
    if (*n_words + 1 > words_cap) { bwt_set_error("bwt_pack_reads: %lld words needed (+1 pad)", (long long)*n_words); return BWT_ERR_NOSPACE; }
    std::vector<std::vector<uint32_t>> exc((size_t)threads);
    auto work = [&](int t) {
        const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
        uint32_t *o = words + wbase[(size_t)t];
        for (int64_t r = lo; r < hi; r++) {
            const int64_t l = offsets[r + 1] - offsets[r];
            lens[r] = (uint16_t)(l > 65535 ? 65535 : l);
            if (l > BWT_PACK_MAX) { exc[(size_t)t].push_back((uint32_t)r); continue; }
            if (!pack_one(reads + offsets[r], (int)l, o)) exc[(size_t)t].push_back((uint32_t)r);
            o += (l + 15) >> 4;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    words[*n_words] = 0;                                   // the pad word every packed buffer carries
    int64_t ne = 0, nb = 0;
    for (auto &v : exc) for (uint32_t r : v) { ne++; nb += offsets[r + 1] - offsets[r]; }
    *n_exc = ne; *exc_bytes = nb;
    if (ne > exc_cap || nb > exc_ascii_cap || (ne > 0 && (!exc_idx || !exc_off || !exc_ascii))) {
        bwt_set_error("bwt_pack_reads: %lld exception reads with %lld bytes need room", (long long)ne, (long long)nb);
        return BWT_ERR_NOSPACE;
    }
    int64_t k = 0, b = 0;
    if (exc_off) exc_off[0] = 0;
    for (auto &v : exc)
        for (uint32_t r : v) {
            const int64_t l = offsets[r + 1] - offsets[r];
            exc_idx[k] = r;
            memcpy(exc_ascii + b, reads + offsets[r], (size_t)l);
            b += l;
            exc_off[++k] = b;
        }
    return BWT_OK;
}

extern "C" int bwt_pack_index(const uint16_t *lens, int64_t n, int64_t *word_index, uint16_t *maxlen_index) {
    if (!lens || n < 0 || !word_index || !maxlen_index) { bwt_set_error("bwt_pack_index: bad argument"); return BWT_ERR_ARG; }
    int64_t w = 0;
    word_index[0] = 0;
    for (int64_t b = 0, k = 0; b < n; b += BWT_PACK_INDEX_STRIDE, k++) {
        const int64_t e = std::min<int64_t>(n, b + BWT_PACK_INDEX_STRIDE);
        uint16_t m = 0;
        for (int64_t r = b; r < e; r++) {
            const uint16_t l = lens[r];
            m = std::max(m, l);
            if (l <= BWT_PACK_MAX) w += (l + 15) >> 4;
        }
        maxlen_index[k] = m;
        word_index[k + 1] = w;
    }
    return BWT_OK;
}
