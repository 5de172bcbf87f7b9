// Candidate extension: sw:sw/2 (sw.erl:47-160) for a list of (query, window) pairs.
#pragma once
#include "common.cuh"

struct SwDesc {          // one (query, window) pair: read `read` of the batch against bytes [r_off, r_off+lr) of the window source
    int64_t r_off;
    uint32_t read;
    int32_t lq;
    int32_t lr;
    uint32_t pad;
};

struct SwOut {           // 16 B per pair
    int32_t score;       // trunc(Vmax)
    int32_t status;      // BWT_HIT_*
    int32_t cigar_len;
    uint32_t cigar_off;  // into the temporary CIGAR arena
};

struct SwCounters {
    unsigned long long sw_calls, cells;
    unsigned long long shortcut_calls, shortcut_cells;   // pairs settled by sw_classify_kernel without the DP (included above)
    uint32_t n_fast, n_wide, n_general, n_ovf;   // list cursors
    uint32_t n_wfa, n_gen2;                      // pairs the quick test left to the wavefront rule; pairs sw_double_kernel handed on to the warp kernel
    uint32_t n_big, gen_cursor;                  // general pairs too large for the warp kernel's slots; its work cursor
    uint32_t cig_used, cig_overflow; // temporary CIGAR arena
    uint32_t general_done, pad;
    unsigned long long band_calls, band_cells;   // pairs settled by sw_band_kernel and the cells sw:sw/2 fills for them (both INCLUDED in sw_calls / cells)
    unsigned long long band_cells_filled;        // cells of the bands it filled instead
    uint32_t n_fast2, n_fast3;                   // cursors of the lists sw_band_kernel (first / second pass) leaves to sw_fast_kernel
    // The temporary CIGAR arena is handed out from SW_CIG_SHARDS regions, each with its cursor on its own 128-byte line
    // (chosen by block): millions of appends to ONE cursor serialise in a single L2 slice at ~1.75 ns each.
    uint32_t cig_sub[64 * 32];
};
constexpr int SW_CIG_SHARDS = 64;

#ifdef __CUDACC__
// `len` bytes of the temporary CIGAR arena -> offset; ok = false (and the overflow flag raised) when the block's region is full
__device__ __forceinline__ uint32_t sw_cig_alloc(SwCounters *ctr, uint32_t cig_cap, uint32_t len, bool &ok) {
    const uint32_t r = blockIdx.x % SW_CIG_SHARDS, cap_r = cig_cap / SW_CIG_SHARDS;
    const uint32_t local = atomicAdd(&ctr->cig_sub[32 * r], len);
    ok = (uint64_t)local + len <= cap_r;
    if (!ok) atomicExch(&ctr->cig_overflow, 1u);
    return r * cap_r + (ok ? local : 0u);
}
#endif

constexpr int32_t SW_FULL_M = 0x100;  // flag on SwOut.status (internal): the CIGAR is "<Lq>M" (pairs settled without the table)
constexpr int SW_R = 10;             // rows per register strip; ten 3-bit direction codes per 32-bit word
constexpr int SW_BLOCK = 128;
constexpr int SW_SHORTCUT_MAX_D = 16; // widest window slack (Lr - Lq) the no-DP shortcut looks at
constexpr int SW_FAST_MAX_LQ = 400;  // query codes of the fast kernel: 25 words per lane; int16 keys are safe up to ~2000

// Device-resident state of the extension stage (owned by a batch).
struct SwStage {
    int device = 0;
    uint32_t cap = 0;                // pairs
    SwDesc *d_desc = nullptr;
    SwOut *d_out = nullptr;
    uint32_t *d_list_fast = nullptr, *d_list_wide = nullptr, *d_list_gen = nullptr, *d_list_ovf = nullptr, *d_list_wfa = nullptr;
    SwCounters *d_ctr = nullptr, *h_ctr = nullptr;     // h_ctr pinned
    char *d_cig_tmp = nullptr;
    uint32_t cig_cap = 0;
    uint2 *d_dirs = nullptr;          // direction scratch: one uint2 per (strip, column, thread)
    size_t dirs_words = 0;
    uint32_t *d_list_big = nullptr;
    uint32_t *d_list_gen2 = nullptr;   // what sw_double_kernel leaves to the warp-per-pair kernel
    uint8_t *d_swd_scratch = nullptr;  // slices of sw_double_kernel
    size_t swd_scratch_bytes = 0;
    uint32_t *d_list_sorted = nullptr; // d_list_fast ordered by (Lq, Lr)
    uint32_t *d_list_fast2 = nullptr, *d_list_fast3 = nullptr;  // what sw_band_kernel's first / second pass leaves
    uint32_t *d_band_scratch = nullptr; // direction words of sw_band_kernel
    size_t band_scratch_words = 0;
    uint32_t *d_sort_bins = nullptr;   // its counting-sort bins
    size_t sort_bins_cap = 0;
    uint32_t *d_gw_scratch = nullptr; // slots of the warp-per-pair general kernel
    size_t gw_scratch_words = 0;
    uint8_t *d_gen_scratch = nullptr;
    size_t gen_scratch_bytes = 0;
    uint64_t *d_gen_off = nullptr;
    size_t gen_off_cap = 0;
    int2 *d_gen_dims = nullptr;
    uint32_t launches = 0;
    bool small = false;               // worker-sized batch: DP pairs go to the warp-per-pair kernel, fast tiers not launched
    unsigned long long sw_calls = 0, cells = 0, n_general = 0, shortcut_calls = 0, shortcut_cells = 0;
    unsigned long long band_calls = 0, band_cells = 0, band_cells_filled = 0;
    int cfg_max_lq = -1, cfg_sms = 0, cfg_per_sm[3] = {0, 0, 0};   // launch configuration of the last run (host queries are not free)
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // around sw_fast_kernel
    float ms_fast_kernel = 0;
};

// thread-per-pair double-precision kernel (swd.cu): the throughput path of pairs with 'N' / IUPAC bytes
struct SwdPlan { int cap_cols, max_strips, blocks; uint64_t off_bdir, off_dirs, scratch_bytes; };
int sw_double_plan(int max_lq, int sms, uint32_t n_cap, SwdPlan *pl);
int launch_sw_double(const SwdPlan &pl, const DevReads &R, const uint8_t *d_rbase, const SwDesc *d_desc, const uint32_t *d_list,
                     const uint32_t *d_n_list, uint32_t *d_next, uint32_t *d_n_next, uint8_t *d_scratch, SwOut *d_out, SwCounters *d_ctr,
                     char *d_cig, uint32_t cig_cap, cudaStream_t stream);

// thread-per-pair banded table with a certificate (swb.cu): plain pairs whose best alignment has a gap
struct SwbPlan { int rows_cap, blocks, p_short, p_second, p_prev; uint64_t scratch_words; };
int sw_band_plan(int max_lq, int sms, uint32_t n_cap, SwbPlan *pl);
int launch_sw_band(const SwbPlan &pl, const DevReads &R, const uint32_t *d_ref2, const SwDesc *d_desc, const uint32_t *d_list,
                   const uint32_t *d_n_list, uint32_t *d_next, uint32_t *d_n_next, uint32_t *d_scratch, SwOut *d_out,
                   SwCounters *d_ctr, char *d_cig, uint32_t cig_cap, cudaStream_t stream);

// warp-per-pair general kernel (swg.cu)
struct SwGwPlan { int wc_cap; uint64_t slot_words, scratch_words; size_t smem; int blocks; };
int sw_general_warp_plan(int max_lq, int sms, SwGwPlan *pl);
int launch_sw_general_warp(const SwGwPlan &pl, const DevReads &R, const uint8_t *d_rbase, const SwDesc *d_desc, const uint32_t *d_list,
                           const uint32_t *d_n_list, uint32_t first, uint32_t *d_cursor, uint32_t *d_list_big, uint32_t *d_scratch, SwOut *d_out, SwCounters *d_ctr,
                           char *d_cig, uint32_t cig_cap, cudaStream_t stream);

int sw_small_limit();
int sw_stage_reserve(SwStage &st, uint32_t pairs);
void sw_stage_free(SwStage &st);
// The extension of pairs [0,*d_n): d_desc must be filled, d_out[k].status preset to BWT_HIT_WINDOW_OOR for pairs
// to skip and anything else for pairs to align.  max_lq = longest query (sizes the shared-memory tier), n_cap >= *d_n.
// R: the batch's reads; d_rbase / d_ref2 / d_refx: the window source as ASCII, 2-bit words and non-ACGT flags.
int sw_stage_enqueue(SwStage &st, const DevReads &R, const uint8_t *d_rbase, const uint32_t *d_ref2, const uint32_t *d_refx,
                     const uint32_t *d_n, uint32_t n_cap, int max_lq, cudaStream_t stream);
int sw_stage_finish(SwStage &st, const DevReads &R, const uint8_t *d_rbase, const uint32_t *d_ref2, const uint32_t *d_refx,
                    uint32_t n_cap, int max_lq, cudaStream_t stream, bool *redo_emit);
int sw_stage_run(SwStage &st, const DevReads &R, const uint8_t *d_rbase, const uint32_t *d_ref2, const uint32_t *d_refx,
                 const uint32_t *d_n, uint32_t n, int max_lq, cudaStream_t stream);
