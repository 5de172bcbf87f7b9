/*
 * bwt_b200.h -- C ABI of the B200-native read-alignment hot path of cloudozer/BWT.
 *
 * This is the drop-in boundary: the Erlang host (worker_bwt) reaches these entry points through the
 * thin NIF in c_src/bwt_nif.c (INTEGRATION.md shows the binding); tests and bench.py reach the very same
 * symbols through ctypes.  Plain pointers and sizes only.  Every function returns BWT_OK (0) or a
 * negative bwt_status; bwt_last_error() gives the message of the calling thread's last failure.
 * Nothing here aborts the process: what is "let it crash" in the reference (case_clause on a bad read
 * character, badmatch on an out-of-range window) comes back as a per-read / per-hit status.
 *
 * There is no CPU fallback behind any of these calls: they need a CUDA device (sm_100a).
 *
 * Citations (file:line) are into the reference tree, cloudozer/BWT.
 */
#ifndef BWT_B200_H
#define BWT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BWT_B200_ABI_VERSION 5   /* 3: packed reads in, compact hits out, bwt_index_info.fused, bwt_sample_reads_edits; 4: bwt_fs_stats.settled; 5: bwt_stats.sw_band* */

typedef enum bwt_status {
    BWT_OK = 0,
    BWT_ERR_ARG = -1,      /* bad argument */
    BWT_ERR_IO = -2,       /* file could not be read / written */
    BWT_ERR_FORMAT = -3,   /* not a {Meta,FM} external term as written by fm_index:make_index/4 */
    BWT_ERR_CUDA = -4,     /* CUDA runtime error (message in bwt_last_error) */
    BWT_ERR_NOMEM = -5,    /* host or device allocation failed */
    BWT_ERR_NOSPACE = -6,  /* caller-provided output capacity too small; the needed sizes are reported */
    BWT_ERR_NODEVICE = -7, /* no usable CUDA device: this library has no CPU path */
    BWT_ERR_LIMIT = -8     /* input beyond a documented limit (n >= 2^32-2^20 rows, read longer than 1031) */
} bwt_status;

/* per-search status: the three results of bwa:find_seeds/6 (bwa.erl:12-45) + the crash case */
enum {
    BWT_SEEDS = 0,          /* a seed list: rows Sp..Ep, N >= 8 characters matched */
    BWT_NO_SEEDS = 1,       /* no_seeds  (fmi.erl:89-91) */
    BWT_TOO_MANY_SEEDS = 2, /* too_many_seeds (bwa.erl:31,33,44) */
    BWT_SEARCH_CRASH = 3    /* the reference raises (case_clause on the last character, bwa.erl:15-24;
                               element(0,FM) after stepping through the '$' row) */
};

/* per-read status of sga / align */
enum {
    BWT_READ_OK = 0,
    BWT_READ_CRASH = 1,     /* some search of this read would crash the reference worker */
    BWT_READ_LIMIT = 2      /* read longer than BWT_MAX_READ_LEN or more seed ends than BWT_MAX_SEEDS_PER_READ */
};

/* per-hit status of sw / align */
enum {
    BWT_HIT_MATCH = 0,      /* sw:sw/2 returned {Score, Cigar} */
    BWT_HIT_NO_MATCH = 1,   /* sw:sw/2 returned no_match (row threshold, sw.erl:69-71) */
    BWT_HIT_WINDOW_OOR = 2  /* window outside the padded reference: badmatch at worker_bwt.erl:89 */
};

/* per-query status of the bounded-edit search (bwt_fs_find_seeds_batch) */
enum {
    BWT_FS_OK = 0,
    BWT_FS_SHORT = 1,       /* Qsec shorter than Len + Defects: the fs.erl that seeds:generate_fs/3 writes has no usable
                               clause or does not compile (a clause with fewer than Len arguments, seeds.erl:51-55) */
    BWT_FS_LIMIT = 2        /* more than max_raw_hits/2 distinct matching windows, or the search stack outgrew its HBM spill area */
};

#define BWT_MAX_READ_LEN 1031          /* 128 subsequences of sga:sga/6 */
#define BWT_MAX_SEEDS_PER_READ (1 << 22)

typedef struct bwt_index bwt_index;   /* opaque: one FM-index partition + its padded .ref, resident in HBM */
typedef struct bwt_batch bwt_batch;   /* opaque: device/pinned buffers and a stream for one batch in flight */

typedef struct bwt_index_info {
    uint64_t n_rows;        /* size(FM), including the '$' row */
    uint32_t pointers[4];   /* {Pc,Pg,Pt,Last}, 1-based (fmi.erl:66-84) */
    int64_t shift;          /* {shift,S} of the meta proplist (fm_index.erl:139) */
    uint64_t ref_len;       /* bytes of the .ref chunk (unpadded) */
    uint64_t dollar_row;    /* 1-based row whose L is '$' */
    uint64_t device_bytes;  /* HBM held by this handle */
    int32_t device;
    int32_t kmer_k;         /* depth of the deepest k-mer interval table (8..15) */
    int32_t fused;          /* 1: the depth-15 table is the fused one (64 B per 15-mer: interval + first 14 SA values) */
    int32_t reserved;
} bwt_index_info;

/* One candidate of one read: an element of worker_bwt's Cigars fold (worker_bwt.erl:84-97), plus the
 * clusters whose sw:sw/2 said no_match (status != 0), which the reference drops.  40 bytes. */
typedef struct bwt_hit {
    int64_t pos;        /* S - Ref_len + Shift  (worker_bwt.erl:94) */
    int64_t s;          /* sga's S: 0-based end (exclusive) of the read in the chunk (sga.erl:46-64) */
    int64_t cigar_off;  /* byte offset of the CIGAR string in the arena (no terminator) */
    int32_t d;          /* sga's D: window slack; Ref_len = Lq + D */
    int32_t score;      /* trunc(Vmax) (sw.erl:123) */
    int32_t status;     /* BWT_HIT_* */
    int32_t cigar_len;
} bwt_hit;

/* The same candidate in 16 bytes, for bulk results (a third of the bytes over PCIe / NVLink): Pos = s - (Lq + D) + Shift is
 * left to the consumer, and a CIGAR that is exactly "<Lq>M" is flagged instead of stored. */
typedef struct bwt_hit16 {
    uint32_t s;           /* sga's S */
    uint32_t d_flags;     /* bits 0-27: sga's D; bits 28-29: BWT_HIT_* status; bit 30: BWT_HIT16_FULL_M */
    uint32_t cigar_off;   /* byte offset in the CIGAR arena (0 when the CIGAR is not stored) */
    int16_t score;        /* trunc(Vmax) */
    uint16_t cigar_len;   /* 0 when the CIGAR is not stored */
} bwt_hit16;
#define BWT_HIT16_D(h) ((h).d_flags & 0x0FFFFFFFu)
#define BWT_HIT16_STATUS(h) (((h).d_flags >> 28) & 3u)
#define BWT_HIT16_FULL_M 0x40000000u   /* the CIGAR is "<Lq>M" (Lq = the read's length) and takes no arena bytes */

/* Reads as the kernels take them: 2 bits per base (A0 C1 G2 T3), read r REVERSED (its last base in bits 0-1 of its first
 * word), ceil(len/16) words per read, reads one after the other, one zero pad word at the very end.  A read that holds a
 * byte outside ACGT (sga:skip_Ns/1, the 1.6 of bwt.hrl:15 and the case_clause of bwa.erl:15-24 need the bytes), or is
 * longer than 1040 bases (then it has no words), is listed in exc_idx and brings its ASCII text.  bwt_pack_reads fills
 * one of these from an ASCII batch. */
typedef struct bwt_packed_reads {
    int64_t n;
    int32_t fixed_len;          /* > 0: every read has this length and lens is unused */
    int32_t reserved;
    const uint16_t *lens;       /* n lengths (capped at 65535) */
    const uint32_t *words;      /* n_words + 1 words */
    int64_t n_words;
    int64_t n_exc;              /* flagged reads: */
    const uint32_t *exc_idx;    /*   their indices, ascending */
    const int64_t *exc_off;     /*   n_exc + 1 offsets into exc_ascii */
    const uint8_t *exc_ascii;   /*   their text */
    /* optional index for cutting the batch into chunks without scanning the lengths (NULL / 0: scanned when needed):
     * word_index[k] = words before read k * BWT_PACK_INDEX_STRIDE, k = 0 .. ceil(n / stride); maxlen_index[k] = longest read
     * of block k */
    const int64_t *word_index;
    const uint16_t *maxlen_index;
} bwt_packed_reads;
#define BWT_PACK_INDEX_STRIDE 65536

typedef struct bwt_stats {   /* work counters of the last run (definitions: DESIGN.md "units") */
    uint64_t reads, searches, lf_steps, hbm_lf_steps, located_rows, clusters, sw_calls, sw_cells, sw_general;
    float ms_total, ms_seed, ms_sw, ms_emit, ms_h2d, ms_d2h; /* CUDA-event times of the stages on the batch's stream */
    float ms_sga_kernel;     /* sga_kernel alone: events directly around its launch */
    float ms_sw_kernel;      /* sw_fast_kernel alone */
    uint32_t launches;       /* kernels of this library launched by the last run */
    uint32_t reserved;
    uint64_t sw_shortcut, sw_shortcut_cells; /* pairs (and their cells, both INCLUDED in sw_calls / sw_cells) whose sw:sw/2
                                                result was proven without filling the table: a perfect or one-mismatch
                                                diagonal with no competing deletion (sw.cu, sw_shortcut) */
    uint64_t sw_band, sw_band_cells;         /* pairs (and the cells sw:sw/2 fills for them, both INCLUDED in sw_calls / sw_cells)
                                                settled from a band of the table plus a certificate (swb.cu) */
    uint64_t sw_band_filled;                 /* cells of the bands filled instead */
} bwt_stats;

const char *bwt_last_error(void);
int bwt_abi_version(void);
int bwt_device_count(void);

/* ---- index lifetime -------------------------------------------------------------------------------- */

/* fm_index:get_index(Chrom,J) (fm_index.erl:154-156) + the .ref load and 200-'N' padding of
 * worker_bwt.erl:48-55: parse <..>_pJ.fm (Erlang external term), upload, build the device layout. */
int bwt_open(const char *fm_path, const char *ref_path, int device, bwt_index **out);

/* Same, from arrays already in host memory (L: n ASCII bytes over "$ACGT"; SA: n 0-based suffix starts). */
int bwt_open_arrays(const uint8_t *L, uint64_t n, const uint32_t *SA, const uint32_t pointers[4], int64_t shift,
                    const uint8_t *ref, uint64_t ref_len, int device, bwt_index **out);

/* fm_index:make_index/4 (fm_index.erl:127-141) on the device: chunk = bytes of the .ref (IUPAC allowed,
 * mapped as fm_index.erl:131-135; 'N' is rejected because the reference substitutes random:uniform(4)).
 * Suffix array by radix sort of 2-bit packed prefixes; any correct SA equals dsa:dsa/1's (dsa.erl:15-48). */
int bwt_build(const uint8_t *chunk, uint64_t len, int64_t shift, int device, bwt_index **out);
/* The same for chunks that hold 'N': each one is replaced in the indexed text (not in the .ref) by
 * element(random:uniform(4), {$A,$C,$G,$T}) as fm_index.erl:131 does, drawn from OTP's `random` generator (AS183) whose
 * state {A1,A2,A3} the caller carries from chunk to chunk of one make_indices run (in/out; a fresh Erlang process starts at
 * {3172, 9814, 20125}).  as183_state = NULL: 'N' is rejected, as bwt_build does. */
int bwt_build_n(const uint8_t *chunk, uint64_t len, int64_t shift, int device, uint32_t as183_state[3], bwt_index **out);

/* Synthetic reference for the benchmark configs: len i.i.d. uniform ACGT bases from a counter-based
 * generator (seed), generated on the device, then indexed like bwt_build. */
int bwt_build_synthetic(uint64_t len, uint64_t seed, int64_t shift, int device, bwt_index **out);

/* Write <fm_path> = term_to_binary({Meta,FM}) and <ref_path> exactly as fm_index:make_index/4 does. */
int bwt_save(const bwt_index *h, const char *fm_path, const char *ref_path);

/* Copy the index columns back to the host (any pointer may be NULL): L[n], SA[n], ref[ref_len]. */
int bwt_export(const bwt_index *h, uint8_t *L, uint32_t *SA, uint8_t *ref);

int bwt_info(const bwt_index *h, bwt_index_info *info);

/* Host only (no GPU needed): rows, pointers and shift of a <Chrom>_pJ.fm as fm_index:get_index/2 would return them
 * (fm_index.erl:154-156).  Atoms in any of the external-format encodings (ATOM_EXT of OTP 17, the UTF-8 forms of OTP 26+). */
int bwt_fm_inspect(const char *fm_path, uint64_t *n_rows, uint32_t pointers[4], int64_t *shift);

/* Depth of the k-mer interval table: entry = state of bwa:find_seeds/6 (bwa.erl:12-45) after the last k characters
 * of a subsequence, so those k-1 LF steps cost one lookup.  0 or 8 = the 4^8 table only; 9..15 adds a 4^k-entry
 * table (8 bytes each; k = 15 = MAX_LEN makes a whole search one lookup, 8.6 GB).  Chosen automatically at open
 * (largest k with 4^k <= 4 * rows); results never depend on it.  Not thread-safe against running batches.
 * Besides this table the handle lazily builds (under its own lock, on first use) one table of exactly the depth of the
 * short subsequence at the front of a batch's reads (8 + Lq mod 8 characters; 134 MB at depth 12), so that those
 * searches need no Occ walk either; BWT_KMER_TAIL=0 in the environment disables it. */
int bwt_set_kmer_depth(bwt_index *h, int k);
int bwt_close(bwt_index *h);

/* ---- one-shot calls with HOST buffers (copies inside) ------------------------------------------------ */

/* bwa:find_seeds/6 (bwa.erl:12-45) for n subsequences: subs = concatenated ASCII, offsets[n+1].
 * Outputs (each n long, any may be NULL): status (BWT_SEEDS..), final Sp/Ep (1-based rows, as the
 * reference numbers them) and the number of characters matched.  The seed ends of search k are
 * SA[Ep]+N, ..., SA[Sp]+N (fmi.erl:89-100): fetch them with bwt_locate. */
int bwt_search_batch(bwt_index *h, const uint8_t *subs, const int64_t *offsets, int64_t n,
                     int32_t *status, uint32_t *sp, uint32_t *ep, int32_t *nmatched);

/* fmi:get_seed_ends/4 payload: SA of rows first..first+count-1 (1-based) */
int bwt_locate(bwt_index *h, uint32_t first_row, uint32_t count, uint32_t *sa_out);

/* sga:sga/6 (sga.erl:17-64) for n reads.  cl_offsets[n+1]; clusters of read r are
 * S[cl_offsets[r]..cl_offsets[r+1]) / D[..] in the reference's order (descending S).
 * *total = number of clusters; BWT_ERR_NOSPACE if cap is too small (then *total = needed). */
int bwt_sga_batch(bwt_index *h, const uint8_t *reads, const int64_t *offsets, int64_t n,
                  int32_t *read_status, int64_t *cl_offsets, int64_t *S, int32_t *D, int64_t cap, int64_t *total);

/* sw:sw/2 (sw.erl:47-160) for n (query, window) pairs.  status BWT_HIT_MATCH / BWT_HIT_NO_MATCH,
 * score = trunc(Vmax), CIGAR bytes of pair k at cigars[cigar_off[k] .. +cigar_len[k]). */
int bwt_sw_batch(int device, const uint8_t *q, const int64_t *q_offsets, const uint8_t *r, const int64_t *r_offsets,
                 int64_t n, int32_t *status, int32_t *score, int64_t *cigar_off, int32_t *cigar_len,
                 char *cigars, int64_t cigar_cap, int64_t *cigar_total);

/* The worker's inner loop for one list of reads (worker_bwt.erl:73-108), minus the top-1 pick that
 * stays in Erlang (worker_bwt.erl:99-106): for read r, hits[hit_offsets[r]..hit_offsets[r+1]) are ALL
 * sga clusters in ascending position; those with status BWT_HIT_MATCH are exactly the reference's Cigars
 * list, in its order.  BWT_ERR_NOSPACE (with *n_hits / *cigar_bytes = needed) if a capacity is too small.
 * The handle keeps one pipelined aligner (streams, events, device buffers sized to the largest call so far) between
 * calls, so that worker-sized calls (200-read lists) do not pay its construction every time; a concurrent second call
 * builds a temporary one.  bwt_close releases it. */
int bwt_align_batch(bwt_index *h, const uint8_t *reads, const int64_t *offsets, int64_t n,
                    int32_t *read_status, int64_t *hit_offsets, bwt_hit *hits, int64_t hits_cap, int64_t *n_hits,
                    char *cigars, int64_t cigar_cap, int64_t *cigar_bytes);

/* ---- resident batches: upload / run / download separately (what bench.py times) --------------------- */

int bwt_batch_create(bwt_index *h, int64_t max_reads, int64_t max_read_bytes, bwt_batch **out);
int bwt_batch_destroy(bwt_batch *b);
/* host -> HBM (pinned staging inside when the caller's memory is pageable) */
int bwt_batch_upload(bwt_batch *b, const uint8_t *reads, const int64_t *offsets, int64_t n);
/* all kernels of the path on the data resident in HBM; results stay in HBM.  Blocks until done. */
int bwt_batch_run(bwt_batch *b);
/* sizes of the last run's result */
int bwt_batch_result_size(bwt_batch *b, int64_t *n_hits, int64_t *cigar_bytes);
/* HBM -> host */
int bwt_batch_download(bwt_batch *b, int32_t *read_status, int64_t *hit_offsets, bwt_hit *hits, int64_t hits_cap,
                       char *cigars, int64_t cigar_cap);
int bwt_batch_stats(bwt_batch *b, bwt_stats *out);

/* The same with packed reads in and compact records out: bwt_batch_upload_packed / bwt_batch_run_compact /
 * bwt_batch_download_compact (uint8 read status, uint32 offsets, bwt_hit16). */
int64_t bwt_pack_words(const int64_t *offsets, int64_t n);     /* words bwt_pack_reads will write (excluding the pad word); -1: bad offsets */
/* ASCII batch -> packed layout on `threads` host threads.  words (capacity words_cap >= bwt_pack_words + 1) and lens[n]
 * always; the flagged reads into exc_idx / exc_off / exc_ascii (BWT_ERR_NOSPACE with *n_exc / *exc_bytes = needed). */
int bwt_pack_reads(const uint8_t *reads, const int64_t *offsets, int64_t n, int threads, uint32_t *words, int64_t words_cap,
                   uint16_t *lens, uint32_t *exc_idx, int64_t exc_cap, int64_t *exc_off, uint8_t *exc_ascii,
                   int64_t exc_ascii_cap, int64_t *n_words, int64_t *n_exc, int64_t *exc_bytes);
/* the chunking index of a packed batch: word_index[n / stride + 2], maxlen_index[n / stride + 1] (stride = BWT_PACK_INDEX_STRIDE) */
int bwt_pack_index(const uint16_t *lens, int64_t n, int64_t *word_index, uint16_t *maxlen_index);
int bwt_batch_upload_packed(bwt_batch *b, const bwt_packed_reads *p);
int bwt_batch_run_compact(bwt_batch *b);
int bwt_batch_download_compact(bwt_batch *b, uint8_t *read_status, uint32_t *hit_offsets, bwt_hit16 *hits, int64_t hits_cap,
                               char *cigars, int64_t cigar_cap);
/* device pointers of the last run's result (wide or compact, as it was run; valid until the next run) */
int bwt_batch_device_result(bwt_batch *b, void **d_hit_offsets, void **d_hits, void **d_cigars);

/* ---- pipelined aligner: bwt_align_batch with persistent buffers and overlapped copies ------------------- */

/* `depth` resident batches of up to chunk_reads reads / chunk_bytes read bytes, each driven by its own host thread
 * and CUDA stream: the H2D copy of one chunk, the kernels of another and the D2H copy of a third overlap.  Same
 * arguments and results as bwt_align_batch; pass pinned host memory for full copy bandwidth.  One run at a time
 * per aligner; several aligners may share one index.  The chunks enter their kernels in chunk order, at most three at
 * a time (BWT_COMPUTE_TOKEN): left to themselves the host threads fall into lock step and the GPU idles while all of
 * them copy; depth 5 is the measured best with 1 M-read chunks. */
typedef struct bwt_aligner bwt_aligner;
int bwt_aligner_create(bwt_index *h, int64_t chunk_reads, int64_t chunk_bytes, int depth, bwt_aligner **out);
int bwt_aligner_run(bwt_aligner *a, const uint8_t *reads, const int64_t *offsets, int64_t n, int32_t *read_status,
                    int64_t *hit_offsets, bwt_hit *hits, int64_t hits_cap, int64_t *n_hits, char *cigars,
                    int64_t cigar_cap, int64_t *cigar_bytes);
/* packed reads in, compact records out: 28 + ~22 bytes per 100 bp read over PCIe instead of 108 + 54 */
int bwt_aligner_run_packed(bwt_aligner *a, const bwt_packed_reads *p, uint8_t *read_status, uint32_t *hit_offsets,
                           bwt_hit16 *hits, int64_t hits_cap, int64_t *n_hits, char *cigars, int64_t cigar_cap,
                           int64_t *cigar_bytes);
int bwt_aligner_stats(bwt_aligner *a, bwt_stats *out);   /* sums over the chunks of the last run */
int bwt_aligner_destroy(bwt_aligner *a);

/* ---- genome-partition mode: the gather of the per-partition records (BASELINE.json configs[4]) ---------------------- */

/* One _pJ partition per GPU (fm_index.erl:113-124 writes them; worker_bwt.erl:48,53 loads one), every GPU aligns ALL reads
 * against its partition, and the records of every partition go to the root over NCCL (NVLink / NVSwitch) straight out of
 * the batches' device-resident compact results.  The reference defines no reduction across partitions (master.erl:151-155
 * prints every cigar it receives): the root gets the partitions' records side by side, in rank order.
 * NCCL is bound at run time (libnccl.so.2); with world = 1 it is never touched.  The 128-byte id is made on one rank
 * (bwt_comm_unique_id) and handed to the others by the host program (the Erlang host's distribution, MPI, a file...). */
typedef struct bwt_comm bwt_comm;
int bwt_comm_unique_id(uint8_t id[128]);
int bwt_comm_create(const uint8_t id[128], int world, int rank, int device, bwt_comm **out);
int bwt_comm_destroy(bwt_comm *c);
/* Collective: every rank passes the batch it ran with bwt_batch_run_compact on the SAME n reads.  On the root, rank p's
 * block lands at status_out[p*n], hit_off_out[p*(n+1)] (rank p's own offsets, from 0), and its hits / CIGAR bytes at the
 * running sums of part_hits / part_cig, which come back on every rank (world entries each).  ms_gather: CUDA-event time
 * of the exchange (sizes all-gather + sends / receives); link_bytes: bytes received by the root over the links.
 * BWT_ERR_NOSPACE on the root when a capacity is too small (the sizes are in part_hits / part_cig). */
int bwt_partition_gather(bwt_comm *c, bwt_batch *b, int root, uint8_t *status_out, uint32_t *hit_off_out, bwt_hit16 *hits_out,
                         int64_t hits_cap, char *cig_out, int64_t cig_cap, int64_t *part_hits, int64_t *part_cig,
                         float *ms_gather, int64_t *link_bytes);

/* ---- bounded-edit search: seeds:generate_fs/3 + the fs:find_seeds/1 it generates (seeds.erl:11-62) -------------- */

typedef struct bwt_fs_stats {
    uint64_t queries, nodes, lf_steps, raw_hits, seeds, spills;  /* search states expanded, LF steps (one child each),
                                                                    distinct matching windows before the walk, seeds kept,
                                                                    stack spills to HBM */
    float ms_kernel;         /* CUDA-event time of fs_search_kernel (all passes) */
    uint32_t launches;
    uint64_t settled;        /* queries answered by the pigeonhole front end (exact piece searches + window verification)
                                without the depth-first search; `nodes` counts their verified candidate windows */
} bwt_fs_stats;

/* For every query Qsec: the list fs:find_seeds(Ref_Seq) returns after seeds:generate_fs(Qsec, Len, Defects), with
 * Ref_Seq = the text of this index partition.  A window of Len characters matches when it fits one of gen_pattern/5's
 * patterns (the first Len+Defects query characters with exactly `Defects` edits: `_` for a character, `_` inserted,
 * character dropped; seeds.erl:51-62); the walk takes the leftmost matching window, jumps Len ahead, repeats
 * (seeds.erl:23-43).  seeds[seed_offsets[r] .. seed_offsets[r+1]) = 1-based window starts N of query r in DESCENDING
 * order, exactly the accumulator the Erlang function returns.  The patterns are never materialised: one warp per
 * query either finds the windows from exact searches of pattern pieces and verifies them (Defects <= 2), or walks the
 * recursion on the FM index (inexact.cu).  The index text is over ACGT (the builder maps IUPAC codes
 * as fm_index.erl:131-135), so a query character outside ACGT can only be covered by a defect.
 * max_raw_hits (<= 0: 1024): per-query scratch for matching windows before the greedy walk; BWT_FS_LIMIT once more
 * than half of it is taken by distinct windows.
 * BWT_ERR_NOSPACE (with *total = needed) when cap is too small; status / seeds / stats may be NULL. */
int bwt_fs_find_seeds_batch(bwt_index *h, const uint8_t *queries, const int64_t *offsets, int64_t n, int32_t len,
                            int32_t defects, int32_t max_raw_hits, int32_t *status, int64_t *seed_offsets,
                            int64_t *seeds, int64_t cap, int64_t *total, bwt_fs_stats *stats);

/* ---- FASTQ ingest (host side; fastq:read_seq/1, fastq.erl:5-18) -------------------------------------------- */

/* Parses up to max_reads records of the FASTQ text [text, text+len) into the packed layout the calls above take:
 * bases appended to reads_out (capacity reads_cap), offsets[0..n], and per record the position/length of its name
 * inside `text` (name = header line minus '@' and minus its last character, exactly as the reference).  Records are
 * returned in FILE order.  final = 0: the buffer may end in the middle of a record, which is then left unconsumed
 * (*consumed = bytes to drop before the next call); final = 1: the text is complete.  No GPU involved. */
int bwt_fastq_parse(const uint8_t *text, int64_t len, int final, int64_t max_reads, uint8_t *reads_out,
                    int64_t reads_cap, int64_t *offsets, int64_t *name_start, int32_t *name_len, int64_t *n_reads,
                    int64_t *consumed);

/* Search-only resident benchmark entry: n fixed-length subsequences already in HBM are searched
 * `repeats` times; *ms = CUDA-event time of all repeats, *lf_steps = LF steps executed per repeat. */
int bwt_search_resident(bwt_index *h, const uint8_t *subs, int64_t n, int32_t sub_len, int32_t repeats,
                        float *ms, uint64_t *lf_steps, uint64_t *hbm_lf_steps);

/* Synthetic reads for the benchmark configs, sampled on the device from the index's own .ref:
 * read r = read_len bases starting at a uniform position (counter-based generator, seed), each base replaced by a
 * different base with probability sub_rate (substitutions only, so every read keeps read_len).  Forward strand.
 * reads_out: host buffer of n*read_len bytes; starts_out (may be NULL): n 0-based chunk positions. */
int bwt_sample_reads(const bwt_index *h, int64_t n, int32_t read_len, uint64_t seed, double sub_rate,
                     uint8_t *reads_out, int64_t *starts_out);

/* The same generator restricted to bases [start, start+len) of that synthetic reference: the text of one _pJ partition
 * (fm_index.erl:113-124 gives start/len; shift = contig shift + Start_J), built where it is needed without holding the
 * whole reference. */
int bwt_build_synthetic_part(uint64_t start, uint64_t len, uint64_t seed, int64_t shift, int device, bwt_index **out);

/* Synthetic reads with the edit model of SURVEY.md 8(d): read r = read_len reference bases from a uniform position, each
 * edited with probability edit_rate -- a share ins_share of the edits inserts one random base after the base, a share
 * del_share deletes the base, the rest substitute it by a different base -- so read lengths vary (read_len + #ins - #del).
 * Afterwards every base turns into 'N' with probability n_rate.  Forward strand.  h != NULL: bases from the index's .ref;
 * h == NULL: from the synthetic reference (synth_len, synth_seed) of bwt_build_synthetic, generated on `device` (every
 * rank of a partitioned run draws the SAME reads without holding the whole text).
 * offsets_out[n+1] always; reads_out (capacity reads_cap) and starts_out may be NULL; *total_bytes = bytes of all reads
 * (BWT_ERR_NOSPACE when reads_cap is too small). */
typedef struct bwt_sample_opts {
    int64_t n;
    int32_t read_len;
    int32_t device;
    uint64_t seed;
    double edit_rate, ins_share, del_share, n_rate;
    uint64_t synth_len, synth_seed;
} bwt_sample_opts;
int bwt_sample_reads_edits(const bwt_index *h, const bwt_sample_opts *o, uint8_t *reads_out, int64_t reads_cap,
                           int64_t *offsets_out, int64_t *starts_out, int64_t *total_bytes);

/* Microbenchmarks that give the roofline denominators measured on the device at hand:
 * random 32-byte-sector gather (two 16-byte loads per sector, L2::64B; GB/s = accesses/s x 32 B) and DPX
 * (VIADDMNMX/VIMNMX3) issue rate (Gop/s). */
int bwt_peak_gather_gbs(int device, uint64_t bytes, double *gbs);
int bwt_peak_dpx_gops(int device, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* BWT_B200_H */
