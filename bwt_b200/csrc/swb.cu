// sw_band_kernel: sw:sw/2 (sw.erl:47-160) of plain-ACGT pairs whose best alignment has a gap, from a band of the table
// plus a certificate (swb_core.cuh holds the argument and the per-row code; tests/test_sw_band_rule.py runs that same
// code on the host against the CPU restatement of sw:sw/2).  One THREAD per pair: the band's three rows of keys (reference rule, affine best,
// affine query-gap state) live in registers, one 2-bit direction code per cell goes to a [row][thread] scratch, the walk
// back and the CIGAR text happen in the same thread.  Pairs the certificate does not cover (a co-optimal path extends a
// gap, penalty beyond the budget, window slack beyond the widest band, reads beyond the scratch rows) are appended to
// `next`, the list sw_fast_kernel fills the whole table for.
//
// Cost: Lq x WB cells of ~23 integer instructions against Lq x Lr cells of 7 in sw_fast_kernel -- 600 cells instead of
// 10,100 for a 100 bp read with one indel.  Launched twice per batch: a small penalty budget (narrow band) on the list
// ordered by shape, then a larger one on what the first pass left (sw.cu, sw_stage_enqueue).
#include <algorithm>
#include <cstdlib>

#include "sw.cuh"
#include "swb_core.cuh"

namespace {

constexpr int SWB_BLOCK = 128;
constexpr int SWB_MIN_BLOCKS = 5;
constexpr int SWB_P_SHORT = 9;
constexpr int SWB_P_SECOND = 15;

struct SwbParams {
    DevReads R;
    const uint32_t *ref2;
    const SwDesc *desc;
    SwOut *out;
    const uint32_t *list, *n_list;
    uint32_t *next, *n_next;
    uint32_t *dirs;               // per block: rows_cap x SWB_BLOCK words
    int rows_cap;
    int p_short;                  // penalty budget of reads up to 128 bp (swb_core.cuh)
    int p_prev;                   // second pass: the budget the pairs of this list were already tried with (0: first pass)
    SwCounters *ctr;
    char *cig;
    uint32_t cig_cap;
};

// the window's 16 bases from base index s (may be negative by a few: the first rows of the band look left of the window)
__device__ __forceinline__ uint32_t window_word(const uint32_t *__restrict__ ref2, int64_t r_off, int s) {
    const int64_t p = r_off + s;
    if (p >= 0) return ref_word(ref2, p);
    return ref_word(ref2, 0) << (2 * (int)(-p));
}

// true = settled (out[k] written)
template <int WB>
__device__ __forceinline__ bool band_pair(const SwbParams &p, const SwDesc &d, uint32_t k, uint32_t *__restrict__ dirs) {
    constexpr int B = SWB_BLOCK;
    const int lq = d.lq, lr = d.lr;
    const int lo = -swb::kk_of(lq, p.p_short);
    const uint32_t *pkr = p.R.pk + __ldg(p.R.woff + d.read);
    swb::Band<WB> s;
    swb::init<WB>(s, lo, lr);
    const int last_full = min(lq, lr - lo - WB + 1);
    // strings by groups of 16 rows: the query word of the group, and the 32 window bases from the group's first cell on --
    // row i = 16g + t + 1 looks at window bases i + lo - 1 + b = (16g + lo) + t + b, one funnel shift away
    uint32_t qw = 0, w0 = 0, w1 = 0;
    auto neq_of = [&](int i) {
        const int t = (i - 1) & 15;
        if (t == 0) {
            qw = read_fwd_word(pkr, lq, (i - 1) >> 4);
            w0 = window_word(p.ref2, d.r_off, i - 1 + lo);
            w1 = window_word(p.ref2, d.r_off, i - 1 + lo + 16);
        }
        const uint32_t qb = (qw >> (2 * t)) & 3u;
        const uint32_t x = __funnelshift_r(w0, w1, 2 * t) ^ (qb * 0x55555555u);
        return (x | (x >> 1)) & 0x55555555u;
    };
    int i = 1;
    for (; i <= lq && i + lo < 1; i++) dirs[(size_t)(i - 1) * B] = swb::row<WB, true>(s, i, lo, lr, neq_of(i));
    for (; i <= last_full; i++) dirs[(size_t)(i - 1) * B] = swb::row<WB, false>(s, i, lo, lr, neq_of(i));
    for (; i <= lq; i++) dirs[(size_t)(i - 1) * B] = swb::row<WB, true>(s, i, lo, lr, neq_of(i));
    int score = 0, b_end = 0;
    if (!swb::finish<WB>(s, lq, lo, lr, p.p_short, &score, &b_end)) return false;
    uint16_t runs[swb::MAX_RUNS];
    // the walk reads the direction words eight rows at a time (independent loads; a walk mostly moves up one row per step)
    uint32_t dw[8];
    int dw_hi = 0;                                               // dw[k] = word of row dw_hi - k; nothing fetched yet
    const int nr = swb::trace<WB>([&](int row) {
        if (row > dw_hi || row <= dw_hi - 8) {
            dw_hi = row;
#pragma unroll
            for (int k = 0; k < 8; k++) dw[k] = row - k >= 1 ? dirs[(size_t)(row - k - 1) * B] : 0u;
        }
        uint32_t v = dw[0];
#pragma unroll
        for (int k = 1; k < 8; k++) if (dw_hi - row == k) v = dw[k];
        return v;
    }, lq, lo, b_end, runs);
    if (nr < 0) return false;
    const int clen = swb::cigar_len(runs, nr);
    bool room;
    const uint32_t off = sw_cig_alloc(p.ctr, p.cig_cap, (uint32_t)clen, room);
    if (room) swb::cigar_write(runs, nr, p.cig + off);
    SwOut o;
    o.score = score; o.status = BWT_HIT_MATCH; o.cigar_len = clen; o.cigar_off = off;
    p.out[k] = o;
    return true;
}

__global__ void __launch_bounds__(SWB_BLOCK, SWB_MIN_BLOCKS) sw_band_kernel(SwbParams p) {
    constexpr int B = SWB_BLOCK;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31;
    uint32_t *dirs = p.dirs + (size_t)blockIdx.x * p.rows_cap * B + threadIdx.x;
    const uint32_t n = *p.n_list;
    unsigned long long cells = 0, bcells = 0;
    uint32_t settled = 0;
    for (uint32_t e0 = blockIdx.x * B; e0 < n; e0 += gridDim.x * B) {
        const uint32_t e = e0 + threadIdx.x;
        const bool todo = e < n;
        uint32_t k = 0;
        SwDesc d{};
        int need = 0;
        if (todo) {
            k = p.list[e]; d = p.desc[k];
            if (d.lq >= 4 && d.lq <= p.rows_cap && d.lr >= d.lq) { need = swb::need_of(d.lq, d.lr, p.p_short); if (need > 16) need = 0; }
            if (p.p_prev && swb::pmax_of(d.lq, p.p_short) <= swb::pmax_of(d.lq, p.p_prev)) need = 0;      // nothing new to try
        }
        // one instantiation per warp: the widest band any of its pairs needs (lists ordered by shape keep warps uniform)
        const int wneed = __reduce_max_sync(FULL, need);
        bool ok = false;
        int wb = 0;
        if (need) {
            if (wneed <= 6) { ok = band_pair<6>(p, d, k, dirs); wb = 6; }
            else if (wneed <= 8) { ok = band_pair<8>(p, d, k, dirs); wb = 8; }
            else if (wneed <= 12) { ok = band_pair<12>(p, d, k, dirs); wb = 12; }
            else { ok = band_pair<16>(p, d, k, dirs); wb = 16; }
        }
        __syncwarp();
        if (ok) { settled++; cells += (unsigned long long)d.lq * (unsigned)d.lr; bcells += (unsigned long long)d.lq * (unsigned)wb; }
        const unsigned left = __ballot_sync(FULL, todo && !ok);
        if (left) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(p.n_next, (uint32_t)__popc(left));
            base = __shfl_sync(FULL, base, 0);
            if (todo && !ok) p.next[base + (uint32_t)__popc(left & ((1u << lane) - 1u))] = k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cells += __shfl_down_sync(FULL, cells, o);
        bcells += __shfl_down_sync(FULL, bcells, o);
        settled += __shfl_down_sync(FULL, settled, o);
    }
    if (lane == 0 && settled) {
        atomicAdd(&p.ctr->sw_calls, (unsigned long long)settled);
        atomicAdd(&p.ctr->cells, cells);
        atomicAdd(&p.ctr->band_calls, (unsigned long long)settled);
        atomicAdd(&p.ctr->band_cells, cells);
        atomicAdd(&p.ctr->band_cells_filled, bcells);
    }
}

}  // namespace

int sw_band_plan(int max_lq, int sms, uint32_t n_cap, SwbPlan *pl) {
    pl->rows_cap = std::min(std::max(max_lq, 4), 256);
    // BWT_SW_BAND_P: penalty budget of reads up to 128 bp in the first pass (9, 12 or 15), BWT_SW_BAND_P2: in a second pass
    // over what the first one left (0 = no second pass); A/B in profiles/r2b8_band_budget.log, r2b9_band_passes.log
    const char *e = getenv("BWT_SW_BAND_P"), *e2 = getenv("BWT_SW_BAND_P2");
    const int pv = e && *e ? atoi(e) : SWB_P_SHORT, pv2 = e2 && *e2 ? atoi(e2) : SWB_P_SECOND;
    pl->p_short = pv <= 9 ? 9 : pv <= 12 ? 12 : 15;
    pl->p_second = pv2 <= pl->p_short ? 0 : pv2 <= 12 ? 12 : 15;
    pl->p_prev = 0;
    const uint64_t want = ((uint64_t)n_cap + SWB_BLOCK - 1) / SWB_BLOCK;
    pl->blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sms * SWB_MIN_BLOCKS));
    pl->scratch_words = (uint64_t)pl->blocks * SWB_BLOCK * pl->rows_cap;
    return BWT_OK;
}

int launch_sw_band(const SwbPlan &pl, const DevReads &R, const uint32_t *d_ref2, const SwDesc *d_desc, const uint32_t *d_list,
                   const uint32_t *d_n_list, uint32_t *d_next, uint32_t *d_n_next, uint32_t *d_scratch, SwOut *d_out,
                   SwCounters *d_ctr, char *d_cig, uint32_t cig_cap, cudaStream_t stream) {
    SwbParams p;
    p.R = R; p.ref2 = d_ref2; p.desc = d_desc; p.out = d_out;
    p.list = d_list; p.n_list = d_n_list; p.next = d_next; p.n_next = d_n_next;
    p.dirs = d_scratch; p.rows_cap = pl.rows_cap; p.p_short = pl.p_short; p.p_prev = pl.p_prev;
    p.ctr = d_ctr; p.cig = d_cig; p.cig_cap = cig_cap;
    sw_band_kernel<<<pl.blocks, SWB_BLOCK, 0, stream>>>(p);
    K_CHECK(stream, "sw_band_kernel");
    return BWT_OK;
}
