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
    const uint2 *ktab;         // interval tables of depth 8 .. min(deep_k, 14), one after the other (ktab_off(d); format: KmerEntry below)
    const uint2 *kdeep;        // the depth-15 table (deep_k == 15 and no fused table), else null
    const uint4 *fused;        // depth 15, 64 B per entry: {Sp, width | n << 28, SA[Sp-1 .. Sp-1+14)}: interval AND its first seed ends
    int deep_k;                // deepest table: 8..15
    const uint8_t *ref;        // .ref padded with 200 'N' on both sides (worker_bwt.erl:53-55)
    const uint32_t *ref2;      // the same, 2 bits per base (A0 C1 G2 T3; anything else as A), 16 bases per word
    const uint32_t *refx;      // bit w set: word w of ref2 holds a byte outside ACGT
    uint64_t ref_len_padded;
    uint32_t n;                // rows
    uint32_t P[4];             // 1-based first row whose F is A (=2), C, G, T
    uint32_t last;
    uint32_t dollar;           // 1-based row whose L is '$'
    int64_t shift;
};

// first entry of the depth-d table inside ktab: 4^8 + 4^9 + .. + 4^(d-1)
__host__ __device__ __forceinline__ uint64_t ktab_off(int d) { return ((1ull << (2 * d)) - 65536ull) / 3ull; }

// ---- reads as the kernels see them -----------------------------------------------------------------
// 2 bits per base, REVERSED: counting j from the END of read r (j = Lq-1-i), base j sits at bits 2(j mod 16) of word
// woff[r] + j/16.  The k-mer table index of the subsequence sga:sga/6 cuts 8k bases before the end is then one funnel
// shift at bit 16k.  Reads that hold a byte outside ACGT (or are longer than BWT_PACK_MAX) are flagged in `exc` and
// take the ASCII paths: their text is in `ascii` (all reads when the batch arrived as ASCII, else only those reads).
#define BWT_PACK_MAX 1040
struct DevReads {
    const uint32_t *pk;
    const uint32_t *woff;      // n+1 word offsets into pk
    const uint32_t *len;       // n lengths in bases
    const uint32_t *exc;       // bit r: read r takes the ASCII path
    const uint8_t *ascii;
    const int64_t *aoff;       // ascii_full: n+1 offsets of all reads, else n_exc+1 offsets of the flagged reads
    const uint32_t *exc_idx;   // !ascii_full: indices of the flagged reads, ascending
    uint32_t n_exc;
    int ascii_full;
    int64_t aoff_base;         // subtracted from every aoff entry (the caller's offsets go up as they are)
    uint32_t exc_base;         // subtracted from every exc_idx entry (a chunk of a larger packed batch)
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
    std::mutex *tail_mu = nullptr;                 // guards the aligner cache
    void *d_occ = nullptr, *d_sa = nullptr, *d_ktab = nullptr, *d_ref = nullptr, *d_kdeep = nullptr, *d_fused = nullptr;
    void *d_ref2 = nullptr, *d_refx = nullptr;
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

// reversed 2-bit groups of a word: group t <-> group 15-t
__device__ __forceinline__ uint32_t rev2(uint32_t x) {
    x = __brev(x);
    return ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
}

__device__ __forceinline__ bool read_is_exc(const DevReads &R, uint32_t r) { return (__ldg(R.exc + (r >> 5)) >> (r & 31)) & 1u; }

// ASCII text of read r (only valid for reads whose text is resident: all of them when ascii_full, else the flagged ones)
__device__ __forceinline__ const uint8_t *read_ascii(const DevReads &R, uint32_t r, int &lq) {
    uint32_t k = r;
    if (!R.ascii_full) {
        uint32_t lo = 0, hi = R.n_exc;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(R.exc_idx + mid) - R.exc_base < r) lo = mid + 1; else hi = mid; }
        k = lo;
    }
    const int64_t o = R.aoff[k];
    const int64_t l = R.aoff[k + 1] - o;
    lq = l > (1 << 30) ? (1 << 30) : (int)l;
    return R.ascii + (o - R.aoff_base);
}

// Bases [16w, 16w+16) of a packed read (pkr = its first word, lq its length) in FORWARD order: base 16w+t at bits 2t,
// zeros ('A') past the end.  One word beyond the read's last may be read (every packed buffer carries a pad word).
__device__ __forceinline__ uint32_t read_fwd_word(const uint32_t *__restrict__ pkr, int lq, int w) {
    const int jlo = lq - 16 - 16 * w;                 // reversed index of base 16w+15
    uint32_t x;
    if (jlo >= 0) x = __funnelshift_r(__ldg(pkr + (jlo >> 4)), __ldg(pkr + (jlo >> 4) + 1), 2 * (jlo & 15));
    else x = __ldg(pkr) << (2 * -jlo);
    return rev2(x);
}
// code of base i (forward position) of a packed read
__device__ __forceinline__ int read_code(const uint32_t *__restrict__ pkr, int lq, int i) {
    const int j = lq - 1 - i;
    return (int)((__ldg(pkr + (j >> 4)) >> (2 * (j & 15))) & 3u);
}

// 16 bases of the packed reference starting at (padded) position p: base p+t at bits 2t
__device__ __forceinline__ uint32_t ref_word(const uint32_t *__restrict__ ref2, int64_t p) {
    const int64_t w = p >> 4;
    return __funnelshift_r(__ldg(ref2 + w), __ldg(ref2 + w + 1), 2 * (int)(p & 15));
}
// true when bases [p, p+len) of the packed reference are all plain ACGT (conservative to 16-base words)
__device__ __forceinline__ bool ref_plain(const uint32_t *__restrict__ refx, int64_t p, int len) {
    if (len <= 0) return true;
    const int64_t b0 = p >> 4, b1 = (p + len - 1) >> 4;
    for (int64_t w = b0 >> 5; w <= (b1 >> 5); w++) {
        uint32_t m = __ldg(refx + w);
        if (w == (b0 >> 5)) m &= 0xFFFFFFFFu << (b0 & 31);
        if (w == (b1 >> 5)) m &= 0xFFFFFFFFu >> (31 - (b1 & 31));
        if (m) return false;
    }
    return true;
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
// 2-bit copy + non-ACGT flags of an ASCII buffer on the device (len bytes; reads 16-byte aligned words: d_ascii from
// cudaMalloc with >= 16 bytes of slack): *d_ref2 gets len/16 + 4 words, *d_refx one bit per word.
int pack_ref_device(const uint8_t *d_ascii, uint64_t len, uint32_t **d_ref2, uint32_t **d_refx, cudaStream_t stream);
// Destroys the aligner bwt_align_batch keeps with the index (batch.cu); called by bwt_close.
void aligner_cache_drop(bwt_index *h);
int index_upload_ref_device(const uint8_t *d_ref, uint64_t len, uint8_t **d_out);

// ---- builder (build.cu) -------------------------------------------------------------------------
int build_sa_device(const uint8_t *d_text_acgt, uint64_t len, uint32_t **d_sa_out, uint8_t **d_L_out,
                    uint32_t pointers[4]);
