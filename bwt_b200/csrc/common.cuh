// Shared declarations of libbwt_b200: error plumbing, the device-side index view, small utilities.
// sm_100a only; no fallbacks.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bwt_b200.h"

// ---- reference constants (include/bwt.hrl:1-20, worker_bwt.erl:16-17) ----------------------------
#define BWT_TOLERANCE 20
#define BWT_MIN_LEN 8
#define BWT_MAX_RANGE 45
#define BWT_MAX_LEN 15
#define BWT_REF_EXT 200

// ---- error plumbing -----------------------------------------------------------------------------
void bwt_set_error(const char *fmt, ...);

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            bwt_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
            return BWT_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

// After a kernel launch: catches launch errors; with BWT_DEBUG_SYNC=1 in the environment it also waits for the
// kernel so that a device fault is reported with the name of the kernel that raised it.
bool bwt_debug_sync();
#define K_CHECK(stream, name)                                                                         \
    do {                                                                                              \
        cudaError_t _e = cudaGetLastError();                                                          \
        if (_e == cudaSuccess && bwt_debug_sync()) _e = cudaStreamSynchronize(stream);                \
        if (_e != cudaSuccess) {                                                                      \
            bwt_set_error("%s:%d: kernel %s -> %s", __FILE__, __LINE__, name, cudaGetErrorString(_e)); \
            return BWT_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

#define BWT_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != BWT_OK) return _s;  \
    } while (0)

// ---- device view of one index partition ---------------------------------------------------------
// Occ layout: one 32-byte block (= one DRAM sector) per 64 rows of L:
//   uint4 #0 : occurrences of A,C,G,T in rows before the block   (checkpoint)
//   uint4 #1 : bit planes of the 64 symbols: lo[0],lo[1],hi[0],hi[1]   (code = hi<<1 | lo, A0 C1 G2 T3;
//              the '$' row is stored as A and corrected with dollar_row)
// so that one LF bound = one aligned 32-byte fetch (two LDG.128 to the same sector).
struct DevIndex {
    const uint4 *occ;
    const uint32_t *sa;        // full suffix array, SA[e-1] for 1-based row e (fmi.erl:89-100 reads it un-sampled)
    const uint2 *kmer;         // 4^8 entries, depth-8 interval table (format: KmerEntry below)
    const uint2 *kmer_deep;    // 4^deep_k entries, or null
    const uint2 *kmer_tail;    // ONE table of depth tail_k (8 < tail_k < deep_k), or null: set per launch for the length of
    int tail_k;                // the short subsequence at the front of the batch's reads (8 + Lq mod 8 characters)
    int deep_k;                // 9..15, or 0 when there is no deep table
    const uint8_t *ref;        // .ref padded with 200 'N' on both sides (worker_bwt.erl:53-55)
    uint64_t ref_len_padded;
    uint32_t n;                // rows
    uint32_t P[4];             // 1-based first row whose F is A (=2), C, G, T
    uint32_t last;
    uint32_t dollar;           // 1-based row whose L is '$'
    int64_t shift;
};

// k-mer interval table entry (uint2): the state of bwa:find_seeds/6 after walking the last k characters of a
// subsequence (all plain ACGT; a trailing 'N' counts as 'A', bwa.erl:21-23):
//   x = Sp                                   y = (Ep - Sp) | n << 28
//   n == k : all k characters matched, interval [Sp,Ep]
//   1 <= n < k : se_down said not_found at character n+1 (bwa.erl:37-38); [Sp,Ep] is the interval at that point
//   n == 0 : Ep - Sp does not fit 28 bits: the caller walks from the start instead
constexpr uint32_t KMER_W_MASK = 0x0FFFFFFFu;
__host__ __device__ __forceinline__ uint2 kmer_pack(uint32_t sp, uint32_t ep, int n) {
    uint32_t w = ep - sp;
    if (w >= KMER_W_MASK) return make_uint2(0u, 0u);
    return make_uint2(sp, w | ((uint32_t)n << 28));
}

struct bwt_index {
    DevIndex dv{};
    int device = 0;
    uint64_t n = 0;
    uint64_t ref_len = 0;      // unpadded
    uint32_t pointers[4] = {0, 0, 0, 0};
    int64_t shift = 0;
    uint64_t device_bytes = 0;
    // one pipelined aligner kept for bwt_align_batch (the NIF calls it per 200-read list: creating streams, events and
    // two dozen device buffers per call would cost far more than the alignment); batch.cu owns it
    void *cached_aligner = nullptr;
    bool cached_busy = false;
    int64_t cached_reads = 0, cached_bytes = 0;
    int cached_depth = 0;
    std::mutex *tail_mu = nullptr;                 // guards the lazily built single-depth tables below and the aligner cache
    void *d_tail[BWT_MAX_LEN] = {};                // d_tail[k]: table of depth k, built on first use (index_tail_table)
    void *d_occ = nullptr, *d_sa = nullptr, *d_kmer = nullptr, *d_ref = nullptr, *d_kmer_deep = nullptr;
};

// ASCII -> index code (A0 C1 G2 T3), -1 otherwise
__host__ __device__ __forceinline__ int bwt_code(uint8_t c) {
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

#ifdef __CUDACC__
// Up to 16 ASCII bases at src (any alignment) -> 2-bit codes (c>>1)&3 (A0 C1 T2 G3), base t at bits 2t; ok is cleared
// when a byte is not one of A,C,G,T.  Reads aligned 32-bit words, up to 7 bytes beyond the last base: every buffer the
// kernels see carries 16 bytes of slack.  Four bases are converted and validated at a time.
__device__ __forceinline__ uint32_t pack16_acgt(const uint8_t *src, int cnt, bool &ok) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint32_t *base = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const int sh = (int)(a & 3) * 8;
    uint32_t prev = __ldg(base), out = 0;
#pragma unroll
    for (int t4 = 0; t4 < 4; t4++) {
        const int left = cnt - 4 * t4;
        if (left > 0) {
            const uint32_t next = __ldg(base + t4 + 1);
            uint32_t wv = __funnelshift_r(prev, next, sh);
            prev = next;
            if (left < 4) wv = (wv & ((1u << (8 * left)) - 1u)) | (0x41414141u << (8 * left));   // pad with 'A'
            const uint32_t x = (wv >> 1) & 0x03030303u;
            // the ASCII byte each code stands for, four at a time: A 0x41, C 0x43, G 0x47, T 0x54
            const uint32_t x0 = x & 0x01010101u, x1 = (x >> 1) & 0x01010101u, tt = x1 & ~x0;
            ok &= ((0x40404040u | (x0 << 1) | (x1 << 2) | (tt << 4) | (tt ^ 0x01010101u)) == wv);
            out |= (((x * 0x00041041u) >> 18) & 0xFFu) << (8 * t4);
        }
    }
    return out;
}
#endif

static inline uint64_t div_up_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// ---- device scan (scan.cu) ----------------------------------------------------------------------
// In-place exclusive prefix sum of n uint32 on `stream`; d_total (device, may be null) receives the sum.
// tmp must hold scan_tmp_words(n) uint32.
size_t scan_tmp_words(size_t n);
int exclusive_scan_u32(uint32_t *d_data, size_t n, uint32_t *d_tmp, uint32_t *d_total, cudaStream_t stream,
                       uint32_t *launches);

// ---- index construction helpers (index.cu) ------------------------------------------------------
// Build occ blocks + k-mer table from device-resident ASCII L (n bytes) and SA; takes ownership of d_sa and d_ref_padded.
int index_finish_from_device(bwt_index *h, const uint8_t *d_L, uint64_t n, uint32_t *d_sa, uint8_t *d_ref_padded,
                             uint64_t ref_len, const uint32_t pointers[4], int64_t shift);
int index_select_device(int device);
// Depth-k interval table (8 < k < deep_k) of this index, built on first use and kept; null when it cannot be had.
const uint2 *index_tail_table(bwt_index *h, int k);
// Destroys the aligner bwt_align_batch keeps with the index (batch.cu); called by bwt_close.
void aligner_cache_drop(bwt_index *h);
int index_upload_ref_device(const uint8_t *d_ref, uint64_t len, uint8_t **d_out);

// ---- builder (build.cu) -------------------------------------------------------------------------
int build_sa_device(const uint8_t *d_text_acgt, uint64_t len, uint32_t **d_sa_out, uint8_t **d_L_out,
                    uint32_t pointers[4]);
