// Candidate extension kernels: sw:sw/2 of the reference (sw.erl:47-160), bit-exact.
//
// The recurrence is NOT Gotoh: one matrix of (V, dir) cells; a gap costs -1 when the neighbour's *chosen*
// direction already is that gap and -2 otherwise; ties resolve up_left > up > left (Erlang term order of
// {V,Atom} under lists:max/1, sw.erl:90-94); the query is global, the window is free at both ends; a row
// whose maximum falls below -2*(Lq div 2) + 2*i ends the alignment with no_match (sw.erl:62,69-71).
//
// Fast path (both strings plain ACGT, the common case):  one THREAD per TWO pairs, no inter-thread traffic.
//   * a cell is one 16-bit key  K = 8*V + code,  code: left=1, up=2, up_left=4.  max over the three candidate
//     keys reproduces value-then-direction ordering in one integer max, and the direction dependent gap cost is
//     m + 8*[code bit] with m = K & ~7:
//        diag = m(i-1,j-1) + (8*sigma+4)     up = m(i-1,j) + 8*[K&2] - 14     left = m(i,j-1) + 8*[K&1] - 15
//   * the keys of two independent pairs (same query length) sit in the two halves of a 32-bit register and every
//     operation is a Blackwell DPX / packed-16 instruction: 2 x VIADDMNMX.S16x2 per cell pair for the three-way max
//     with its fused adds, VIADD.16x2 for the left candidate, VIMNMX.S16x2 for the row maximum, one PRMT that
//     looks up both lanes' substitution scores from per-row byte profiles, LOP3 for the code/value split.
//   * the thread sweeps the matrix in strips of SW_R = 10 rows held in registers; only the strip's bottom row
//     (one 32-bit word per column = both pairs) and the 2-bit packed windows live in shared memory, laid out
//     [column][thread] so that a warp's accesses are conflict free.
//   * the ten 3-bit direction codes of one column of a strip are 2 x 15 bits per pair: one uint2 per column holds
//     both pairs' codes, written to a per-thread slice of an HBM/L2 scratch ([column][thread], coalesced) and read
//     back by the traceback.
// General path (any byte, e.g. 'N' = 1.6 or IUPAC codes in the window, or sizes beyond the fast tiers):
//   IEEE doubles added in the reference's order, one thread per pair, rows in HBM scratch.
#include <cstdlib>
#include <type_traits>

#include "sw.cuh"

// -DBWT_KDEBUG builds print and bail out on a violated kernel invariant instead of faulting (debug only).
#ifdef BWT_KDEBUG
#define KD_CHECK(cond, ...) do { if (!(cond)) { printf("KDEBUG %s:%d " #cond " ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); return; } } while (0)
#else
#define KD_CHECK(cond, ...) do { } while (0)
#endif

namespace {

// PRMT in its native form (the __byte_perm intrinsic masks the sign-replicate bit of the selector nibbles off):
// bytes 0-3 come from a, 4-7 from b; a selector nibble with bit 3 set replicates the sign of the chosen byte.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

__device__ __forceinline__ int lane16(uint32_t x, int lane) { return lane ? ((int)x >> 16) : ((int)(x << 16) >> 16); }
__device__ __forceinline__ uint32_t both(int k) { return ((uint32_t)k & 0xFFFFu) * 0x10001u; }

// ---------------------------------------------------------------------------------------------------
// Pairs whose sw:sw/2 result is known without filling the table.
//
// Both strings plain ACGT, window slack Dd = Lr - Lq in 0..SW_SHORTCUT_MAX_D.  Diagonal o pairs query character i with
// window character o + i.  Write a path's score as 2*Lq - penalty: a mismatch costs 3, a skipped window character 2
// (gap open) or 1 (extension), a skipped query character 4 or 3 (the gap and the lost match).  With PLAIN affine gaps
// the cheapest penalty p* and the diagonals on which a path of penalty p* reaches the last row follow from a wavefront
// search (furthest query index per penalty, diagonal and state M / I / D -- the WFA recurrence), a few dozen short
// string comparisons instead of Lq*Lr cells.  The reference's table is not plain affine (a gap is extended at -1 only
// when the neighbour's CHOSEN direction is that gap, sw.erl:88-107), but that rule never scores a path higher than
// plain affine gaps would, and a gapless diagonal scores the same under both.  So if the LEFTMOST diagonal o* on which
// penalty p* reaches the last row does so without a gap (its own mismatch count is p*/3):
//   * the last row's maximum is 2*Lq - p*, and no column left of o* + Lq reaches it, which is the column
//     get_starting_point takes (sw.erl:110-112);
//   * every cell of that diagonal holds exactly the diagonal's prefix score (more would push the end cell past the
//     maximum), its up_left candidate attains it, and ties go up_left > up > left: the traceback is "<Lq>M";
//   * the row threshold -2*(Lq div 2) + 2*i (sw.erl:62,69-71) is never undercut while p* <= 2*(Lq div 2).
// Searched up to penalty SC_PMAX (five mismatches); paths may leave the window's diagonals by up to
// SC_K = (PMAX-1)/3 on either side (column 0 starts with a query-side gap, sw.erl:72-77).  Anything else -- optimal
// paths with a gap, more mismatches, other bytes -- runs the DP.  tests/test_gpu_parity.py runs its SW cases with
// this shortcut on and off, including low-complexity strings built to break it.
// ---------------------------------------------------------------------------------------------------
// Eight lanes per pair: lane t packs bases [16t, 16t+16) of every 128-base chunk (coalesced loads) and owns the
// diagonals t, t+8, t+16 of the wavefront search.
constexpr int SC_LANES = 8;

// reductions over the SC_LANES lanes of a group with FULL-mask shuffles: the four groups of a warp stay converged
// (partial-mask primitives would split the warp into four serialised quarter-warps)
__device__ __forceinline__ unsigned grp_min(unsigned v) {
#pragma unroll
    for (int o = 1; o < SC_LANES; o <<= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
constexpr int SC_PMAX = 15;                         // largest penalty searched: five mismatches
constexpr int SC_K = (SC_PMAX - 1) / 3;             // query characters a path of that penalty can skip
constexpr int SC_ND = 32;                           // diagonals -SC_K .. SW_SHORTCUT_MAX_D + SC_K, padded
constexpr int SC_RING = 5;                          // wavefronts kept: s, s-1, .., s-4
constexpr int SC_MAXQ = 256;                        // longest query the shortcut looks at
constexpr int SC_QW = SC_MAXQ / 16;                 // 2-bit packed query words (16 bases each)
constexpr int SC_RW = (SC_MAXQ + SW_SHORTCUT_MAX_D + 15) / 16 + 2;
static_assert(SW_SHORTCUT_MAX_D + 2 * SC_K + 1 <= SC_ND, "diagonal slots");
static_assert(SC_RING == 5, "the level loop rotates five ring slots");

// Per pair: the wavefronts (furthest query index, -1 = unreachable; states M, I, D) and both strings packed to 2 bits
// per base, so that "extend along diagonal o" compares 16 bases per XOR instead of walking bytes.
struct ScWave {
    int16_t v[SC_RING][3][SC_ND];
    uint32_t qp[SC_QW];
    uint32_t rp[SC_RW];
};

// mismatch flags of query bases [16w, 16w+16) on diagonal o: bit 2b set <=> q[16w+b] != r[o+16w+b]
// (window positions below 0 give garbage in the low flags: callers never look below i = -o)
__device__ __forceinline__ uint32_t sc_diff(const ScWave &wf, int o, int w) {
    const int p = o + 16 * w;
    uint32_t rv;
    if (p >= 0) rv = __funnelshift_r(wf.rp[p >> 4], wf.rp[(p >> 4) + 1], 2 * (p & 15));
    else rv = wf.rp[0] << (2 * -p);
    const uint32_t x = wf.qp[w] ^ rv;
    return (x | (x >> 1)) & 0x55555555u;
}

// furthest i' >= i with q[i..i') == r[o+i..o+i'), i' <= lim
__device__ __forceinline__ int sc_extend(const ScWave &wf, int o, int i, int lim) {
    if (i >= lim) return lim;
    int w = i >> 4;
    uint32_t m = sc_diff(wf, o, w) & (0xFFFFFFFFu << (2 * (i & 15)));
    const int nw = (lim + 15) >> 4;
    while (m == 0u && ++w < nw) m = sc_diff(wf, o, w);
    return m ? min(lim, 16 * w + ((__ffs((int)m) - 1) >> 1)) : lim;
}

// penalty budget of the wavefront search by read length (three mismatches up to 128 bp, four up to 192, five beyond):
// every level costs the whole warp time, and a wider budget also means more diagonals (kk on either side of the window's)
__device__ __forceinline__ int sc_pmax(int lq) {
    return min(lq <= 128 ? 9 : lq <= 192 ? 12 : SC_PMAX, 2 * (lq >> 1));     // the row threshold holds up to 2*(Lq div 2)
}

// 0 = run the DP; otherwise the score of the pair, whose CIGAR is "<Lq>M".  Called by ALL lanes of the warp, converged;
// the SC_LANES lanes of a group pass the same pair (valid = false: no pair) and the group's storage; t = lane index
// inside the group.  Both strings are plain ACGT (the caller checked): pkr = the read's packed words, the window is
// bases [r_off, r_off+lr) of the packed window source ref2.
__device__ __forceinline__ int sw_shortcut(const uint32_t *__restrict__ pkr, int lq, const uint32_t *__restrict__ ref2, int64_t r_off,
                                           int lr, bool valid, int t, unsigned lane, ScWave &wf) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int dd = lr - lq;
    const bool elig = valid && dd >= 0 && dd <= SW_SHORTCUT_MAX_D && lq >= 4 && lq <= SC_MAXQ;
    // both strings as 2-bit words, 16 bases each, forward (lane t: words t, t+8, ..): window padded with 'A' past its end
    if (elig) {
        for (int w = t; w < SC_RW; w += SC_LANES) {
            uint32_t x = 0u;
            const int left = lr - 16 * w;
            if (left > 0) { x = ref_word(ref2, r_off + 16 * w); if (left < 16) x &= (1u << (2 * left)) - 1u; }
            wf.rp[w] = x;
        }
        for (int w = t; 16 * w < lq; w += SC_LANES) wf.qp[w] = read_fwd_word(pkr, lq, w);
    }
    __syncwarp();                                                    // qp / rp visible to the whole group
    (void)lane;
    const int pmax = sc_pmax(lq);
    const int kk = (pmax - 1) / 3;
    const int nd = dd + 2 * kk + 1;                                  // diagonal slot d <-> o = d - kk
    bool open = elig;
    int pstar = -1, omin = 0x7FFF;
    int cur = 0, p1 = 4, p2 = 3, p3 = 2, p4 = 1;                      // ring slots of wavefronts s, s-1, .., s-4
    for (int s = 0; s <= SC_PMAX; s++) {
        const bool act = open && s <= pmax;
        unsigned best = 0x7FFFu;
        if (act) {
            for (int d = t; d < nd; d += SC_LANES) {
                const int o = d - kk;
                int dv = -1, iv = -1, mv = -1;
                if (s == 1) {                                        // nothing costs 1: an empty wavefront
                    wf.v[cur][0][d] = -1; wf.v[cur][1][d] = -1; wf.v[cur][2][d] = -1;
                    continue;
                }
                if (d > 0) {                                         // skip a window character: diagonal o-1 -> o
                    if (s >= 2) dv = wf.v[p2][0][d - 1];
                    if (s >= 1) dv = max(dv, (int)wf.v[p1][2][d - 1]);
                    if (dv >= 0 && (o + dv < 0 || o + dv > lr)) dv = -1;
                }
                if (d + 1 < nd) {                                    // skip a query character: diagonal o+1 -> o
                    if (s >= 4) iv = wf.v[p4][0][d + 1];
                    if (s >= 3) iv = max(iv, (int)wf.v[p3][1][d + 1]);
                    if (iv >= 0) { iv += 1; if (iv > lq || o + iv < 0) iv = -1; }
                }
                if (s == 0) mv = (o >= 0 && o <= lr) ? 0 : -1;       // row 0 is free (sw.erl:72)
                else {
                    const int pm = s >= 3 ? (int)wf.v[p3][0][d] : -1;
                    if (pm >= 0 && pm < lq && o + pm >= 0 && o + pm < lr) mv = pm + 1;   // mismatch
                    mv = max(mv, max(dv, iv));
                }
                if (mv >= 0) mv = sc_extend(wf, o, mv, min(lq, lr - o));
                wf.v[cur][0][d] = (int16_t)mv; wf.v[cur][1][d] = (int16_t)iv; wf.v[cur][2][d] = (int16_t)dv;
                if (mv == lq && o + lq >= 0 && o + lq <= lr) best = min(best, (unsigned)d);
            }
        }
        __syncwarp();
        best = grp_min(best);
        if (act && best != 0x7FFFu) { open = false; pstar = s; omin = (int)best - kk; }
        if (__all_sync(FULL, !open || s >= pmax)) break;
        { const int free_slot = p4; p4 = p3; p3 = p2; p2 = p1; p1 = cur; cur = free_slot; }
    }
    // the leftmost optimal end must be a gapless diagonal of the window: count its mismatches (lane t: words t, t+8, ..)
    const bool check = pstar >= 0 && omin >= 0 && omin <= dd;
    unsigned mism = 0;
    if (check)
        for (int w = t; w < (lq + 15) >> 4; w += SC_LANES) {
            uint32_t f = sc_diff(wf, omin, w);
            if (lq - 16 * w < 16) f &= (1u << (2 * (lq - 16 * w))) - 1u;   // bases past the query's end are padding
            mism += (unsigned)__popc(f);
        }
#pragma unroll
    for (int o = 1; o < SC_LANES; o <<= 1) mism += __shfl_xor_sync(FULL, mism, o);
    if (!check) return 0;
    return 3 * (int)mism == pstar ? 2 * lq - pstar : 0;
}

// ---------------------------------------------------------------------------------------------------
// classification.  Stage 1 (sw_classify_kernel, one thread per pair): pairs the quick test settles, and the routing of
// the rest.  Stage 2 (sw_wfa_kernel, eight lanes per pair): the wavefront rule above on the pairs stage 1 left open.
//
// Quick test = the two cases of the wavefront rule that need no wavefront (both strings plain ACGT, 0 <= Dd <= 16,
// 4 <= Lq <= 256, exactly the rule's own domain):
//   * some diagonal of the window is a PERFECT match: p* = 0, only gapless perfect diagonals cost 0 (a diagonal that
//     starts left of the window begins with a query-side gap, >= 4), and the leftmost of them is o*:  {2*Lq, "<Lq>M"};
//   * Dd = 0 and the one diagonal has ONE mismatch: every other path needs a query-side gap (>= 4 > 3) -- a window-side
//     gap alone leaves the window, column 0 costs 4 -- so p* = 3 on that diagonal:  {2*Lq - 3, "<Lq>M"}.
// Everything else with a chance (more mismatches, Dd > 0 without a perfect diagonal) goes to the wavefront search.
// ---------------------------------------------------------------------------------------------------
// Per-thread work counters of the classification kernels, flushed once per block at the end of the kernel.
struct ScStats { unsigned long long cells = 0; uint32_t settled = 0; };

// The end of a classification round for a whole 256-thread block: list slots for kinds 0 / 1 / 2 / 4 and CIGAR text
// "<Lq>M" + result for kind 3 (settled without the table).  ONE atomic per block and counter: the lanes of a kind take
// consecutive slots (warp ballots, per-warp counts summed in shared memory).  Called by all threads of the block.
__device__ __forceinline__ void sw_route_settle(int kind, uint32_t k, int lq, int lr, int score, SwOut *__restrict__ out, SwCounters *ctr,
                                                uint32_t *__restrict__ list_fast, uint32_t *__restrict__ list_wide,
                                                uint32_t *__restrict__ list_gen, uint32_t *__restrict__ list_wfa,
                                                char *__restrict__ cig, uint32_t cig_cap, ScStats &acc) {
    __shared__ uint32_t s_w[8][5];            // per warp: members of kinds 0, 1, 2, 4 and CIGAR bytes of kind 3
    __shared__ uint32_t s_b[5];               // the block's bases
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = kind == 4 ? 3 : kind;      // list column: kinds 0, 1, 2 -> 0, 1, 2; kind 4 -> 3
    unsigned mine = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, kind != 3 && kind >= 0 && kk == t);
        if (lane == 0) s_w[warp][t] = (uint32_t)__popc(m);
        if (kind != 3 && kind >= 0 && kk == t) mine = m;
    }
    const int clen = kind == 3 ? 2 + (lq >= 10) + (lq >= 100) + (lq >= 1000) + (lq >= 10000) : 0;
    int incl = clen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    if (lane == 31) s_w[warp][4] = (uint32_t)incl;
    __syncthreads();
    if (threadIdx.x < 5) {
        const int t = (int)threadIdx.x;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const uint32_t c = s_w[w][t]; s_w[w][t] = sum; sum += c; }
        uint32_t base = 0;
        if (sum) {
            if (t < 4) base = atomicAdd(t == 0 ? &ctr->n_fast : t == 1 ? &ctr->n_wide : t == 2 ? &ctr->n_general : &ctr->n_wfa, sum);
            else { bool ok; base = sw_cig_alloc(ctr, cig_cap, sum, ok); if (!ok) base = 0xFFFFFFFFu; }
        }
        s_b[t] = base;
    }
    __syncthreads();
    if (kind == 3) {
        acc.settled += 1; acc.cells += (unsigned long long)lq * (unsigned)lr;
        const uint32_t base = s_b[4];
        uint32_t off = 0;
        if (base != 0xFFFFFFFFu) {
            off = base + s_w[warp][4] + (uint32_t)(incl - clen);
            int l = lq;
            for (int t = clen - 2; t >= 0; t--) { cig[off + t] = (char)('0' + l % 10); l /= 10; }
            cig[off + clen - 1] = 'M';
        }
        SwOut o; o.score = score; o.status = BWT_HIT_MATCH | SW_FULL_M; o.cigar_len = clen; o.cigar_off = off;
        out[k] = o;
    } else if (kind >= 0) {
        const uint32_t slot = s_b[kk] + s_w[warp][kk] + (uint32_t)__popc(mine & ((1u << lane) - 1u));
        (kk == 0 ? list_fast : kk == 1 ? list_wide : kk == 2 ? list_gen : list_wfa)[slot] = k;
    }
    __syncthreads();                          // s_w / s_b are reused by the next round
}

// the block's work counters -> global (one atomic per block and counter)
__device__ __forceinline__ void sc_stats_flush(const ScStats &acc, SwCounters *ctr) {
    __shared__ unsigned long long s_cells;
    __shared__ uint32_t s_settled;
    if (threadIdx.x == 0) { s_cells = 0; s_settled = 0; }
    __syncthreads();
    unsigned long long c = acc.cells;
    uint32_t n = acc.settled;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_down_sync(0xFFFFFFFFu, c, o); n += __shfl_down_sync(0xFFFFFFFFu, n, o); }
    if ((threadIdx.x & 31) == 0 && n) { atomicAdd(&s_cells, c); atomicAdd(&s_settled, n); }
    __syncthreads();
    if (threadIdx.x == 0 && s_settled) {
        atomicAdd(&ctr->sw_calls, (unsigned long long)s_settled);
        atomicAdd(&ctr->cells, s_cells);
        atomicAdd(&ctr->shortcut_calls, (unsigned long long)s_settled);
        atomicAdd(&ctr->shortcut_cells, s_cells);
    }
}

__device__ __forceinline__ int sw_dp_kind(const SwDesc &d, int cap_narrow, int cap_wide) {
    if (d.lq <= SW_FAST_MAX_LQ && d.lr + 1 <= cap_narrow) return 0;
    if (d.lq <= SW_FAST_MAX_LQ && d.lr + 1 <= cap_wide) return 1;
    return 2;
}

// use_shortcut: bit 0 = quick test, bit 1 = wavefront rule.
// Lane L of a warp keeps the books of pair k0 + L; the quick test itself runs EIGHT lanes per pair (lane t compares
// words t and t+8 of the query with the window: coalesced loads, no dependent chain per thread): in round i the warp's
// four groups take the pairs of lanes 4i..4i+3, whose descriptors they fetch by shuffle, and hand the verdict back.
__global__ void __launch_bounds__(256) sw_classify_kernel(DevReads R, const uint32_t *__restrict__ ref2, const uint32_t *__restrict__ refx,
                                   const SwDesc *__restrict__ desc, SwOut *__restrict__ out, const uint32_t *__restrict__ n_ptr,
                                   int cap_narrow, int cap_wide, uint32_t *__restrict__ list_fast,
                                   uint32_t *__restrict__ list_wide, uint32_t *__restrict__ list_gen, uint32_t *__restrict__ list_wfa,
                                   SwCounters *ctr, char *__restrict__ cig, uint32_t cig_cap, int use_shortcut) {
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31;
    const int t = (int)(lane & 7u), g = (int)(lane >> 3);
    const uint32_t n = *n_ptr;
    ScStats acc;
    for (uint32_t k0 = blockIdx.x * blockDim.x; k0 < n; k0 += gridDim.x * blockDim.x) {
        const uint32_t k = k0 + threadIdx.x;
        int score = 0, kind = -1;             // kind: 0 narrow, 1 wide, 2 general, 3 settled, 4 wavefront list, -1 nothing to do
        SwDesc d{};
        bool elig = false;
        unsigned long long pk_addr = 0;
        if (k < n && out[k].status != BWT_HIT_WINDOW_OOR) {
            d = desc[k];
            if (d.lq <= 0 || d.lr <= 0) {               // sw(_,_) -> no_match for empty strings (sw.erl:54)
                SwOut o; o.score = 0; o.status = BWT_HIT_NO_MATCH; o.cigar_len = 0; o.cigar_off = 0;
                out[k] = o;
            } else if (read_is_exc(R, d.read) || !ref_plain(refx, d.r_off, d.lr)) {
                kind = 2;                               // 'N' / IUPAC / other bytes on either side: doubles (bwt.hrl:15)
            } else {
                const int dd = d.lr - d.lq;
                elig = dd >= 0 && dd <= SW_SHORTCUT_MAX_D && d.lq >= 4 && d.lq <= SC_MAXQ;
                kind = elig && (use_shortcut & 2) ? 4 : sw_dp_kind(d, cap_narrow, cap_wide);
                if (elig) pk_addr = (unsigned long long)(R.pk + __ldg(R.woff + d.read));
            }
        }
        const bool quick = elig && (use_shortcut & 1);
        if (__any_sync(FULL, quick)) {
#pragma unroll 2
            for (int i = 0; i < 8; i++) {
                const int src = 4 * i + g;
                const bool qk = __shfl_sync(FULL, (int)quick, src) != 0;
                if (!__any_sync(FULL, qk)) continue;
                const uint32_t *pkr = reinterpret_cast<const uint32_t *>(__shfl_sync(FULL, pk_addr, src));
                const long long ro = __shfl_sync(FULL, (long long)d.r_off, src);
                const int lq = __shfl_sync(FULL, d.lq, src), dd = __shfl_sync(FULL, d.lr, src) - lq;
                const int nw = (lq + 15) >> 4;
                const bool h0 = qk && t < nw, h1 = qk && t + 8 < nw;        // this lane's words: t and t + 8
                const uint32_t q0 = h0 ? read_fwd_word(pkr, lq, t) : 0u, q1 = h1 ? read_fwd_word(pkr, lq, t + 8) : 0u;
                const uint32_t m0 = h0 ? (lq - 16 * t < 16 ? (1u << (2 * (lq - 16 * t))) - 1u : 0xFFFFFFFFu) & 0x55555555u : 0u;
                const uint32_t m1 = h1 ? (lq - 16 * (t + 8) < 16 ? (1u << (2 * (lq - 16 * (t + 8)))) - 1u : 0xFFFFFFFFu) & 0x55555555u : 0u;
                int sc = 0, mn = 0x7FFF;                                // mn: fewest mismatches on a diagonal of the window
                const int omax = __reduce_max_sync(FULL, qk ? dd : 0);
                for (int o = 0; o <= omax; o++) {                       // diagonals of the window, leftmost first
                    const bool live = qk && o <= dd && sc == 0;
                    uint32_t f0 = 0u, f1 = 0u;
                    if (live && h0) { const uint32_t x = q0 ^ ref_word(ref2, ro + o + 16 * t); f0 = (x | (x >> 1)) & m0; }
                    if (live && h1) { const uint32_t x = q1 ^ ref_word(ref2, ro + o + 16 * (t + 8)); f1 = (x | (x >> 1)) & m1; }
                    int mism = __popc(f0) + __popc(f1);
                    mism += __shfl_xor_sync(FULL, mism, 1);
                    mism += __shfl_xor_sync(FULL, mism, 2);
                    mism += __shfl_xor_sync(FULL, mism, 4);
                    if (live) {
                        mn = min(mn, mism);
                        if (mism == 0) sc = 2 * lq;                     // the leftmost perfect diagonal
                        else if (dd == 0 && mism == 1) sc = 2 * lq - 3; // one mismatch, no room for a gap
                    }
                }
                // The wavefront rule can only settle a pair on a window diagonal with at most pmax/3 mismatches (its own
                // acceptance test): when every diagonal has more -- the usual picture of a read with an indel, whose
                // diagonals all disagree from the indel on -- the search would run every level and then hand the pair to
                // the DP anyway.  -1 = skip the search.
                if (sc == 0 && 3 * mn > sc_pmax(lq)) sc = -1;
                const int got = __shfl_sync(FULL, sc, 8 * (int)(lane & 3u));
                if ((int)(lane >> 2) == i) score = got;
            }
        }
        if (score > 0) kind = 3;
        else if (score < 0 && kind == 4) kind = sw_dp_kind(d, cap_narrow, cap_wide);
        sw_route_settle(kind, k, d.lq, d.lr, score, out, ctr, list_fast, list_wide, list_gen, list_wfa, cig, cig_cap, acc);
    }
    sc_stats_flush(acc, ctr);
}

// Stage 2: the wavefront rule on the pairs of list_wfa (all plain, all inside the rule's domain).  Lane L of a warp
// keeps the books of list entry base + L (32 pairs per warp: one atomic per warp and list); the rule itself runs
// SC_LANES lanes per pair: in round i the warp's four groups work on the pairs of lanes 4i..4i+3, whose descriptors
// they fetch by shuffle, and hand the verdict back the same way.
__global__ void __launch_bounds__(256) sw_wfa_kernel(DevReads R, const uint32_t *__restrict__ ref2,
                              const SwDesc *__restrict__ desc, SwOut *__restrict__ out, const uint32_t *__restrict__ list_wfa,
                              int cap_narrow, int cap_wide, uint32_t *__restrict__ list_fast,
                              uint32_t *__restrict__ list_wide, uint32_t *__restrict__ list_gen, SwCounters *ctr,
                              char *__restrict__ cig, uint32_t cig_cap) {
    __shared__ ScWave s_wave[256 / SC_LANES];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t n = ctr->n_wfa;
    ScStats acc;
    for (uint32_t e0 = blockIdx.x * blockDim.x; e0 < n; e0 += gridDim.x * blockDim.x) {
        const uint32_t e = e0 + threadIdx.x;
        int score = 0, kind = -1;
        SwDesc d{};
        uint32_t k = 0;
        const bool todo = e < n;
        unsigned long long pk_addr = 0;
        if (todo) { k = list_wfa[e]; d = desc[k]; pk_addr = (unsigned long long)(R.pk + __ldg(R.woff + d.read)); }
        const int t = (int)(lane % SC_LANES), g = (int)(lane / SC_LANES);
#pragma unroll 1
        for (int i = 0; i < 32 / (32 / SC_LANES); i++) {
            const int src = (32 / SC_LANES) * i + g;
            const unsigned long long qa = __shfl_sync(FULL, pk_addr, src);
            const long long ro = __shfl_sync(FULL, (long long)d.r_off, src);
            const int lq = __shfl_sync(FULL, d.lq, src), lr = __shfl_sync(FULL, d.lr, src);
            const bool td = __shfl_sync(FULL, (int)todo, src) != 0;
            if (!__any_sync(FULL, td)) continue;
            const int sc = sw_shortcut(reinterpret_cast<const uint32_t *>(qa), lq, ref2, ro, lr, td, t, lane, s_wave[threadIdx.x / SC_LANES]);
            const int got = __shfl_sync(FULL, sc, SC_LANES * (int)(lane % (32 / SC_LANES)));
            if ((int)(lane / (32 / SC_LANES)) == i) score = got;
        }
        if (todo) kind = score ? 3 : sw_dp_kind(d, cap_narrow, cap_wide);
        sw_route_settle(kind, k, d.lq, d.lr, score, out, ctr, list_fast, list_wide, list_gen, nullptr, cig, cig_cap, acc);
    }
    sc_stats_flush(acc, ctr);
}

// ---------------------------------------------------------------------------------------------------
// fast path
// ---------------------------------------------------------------------------------------------------
struct SwFastParams {
    DevReads R;
    const uint32_t *ref2;
    const SwDesc *desc;
    const uint32_t *list;
    const uint32_t *n_list;   // device counter: entries of list
    SwOut *out;
    uint32_t *list_gen;
    uint32_t *list_ovf;       // pairs to redo with absolute boundary rows
    SwCounters *ctr;
    char *cig;
    uint32_t cig_cap;
    uint2 *dirs;
    int cap_cols;      // columns (Lr+1) the shared row can hold; multiple of 16
    int max_strips;
};

// (a & 0x00080008) | c in ONE LOP3: the instruction takes a single immediate, so c must sit in a register -- and the
// compiler, which knows it is a constant, would rather spend a second LOP3 than a register; sw_opaque hides the value.
// The DP loop is bound by the ALU pipe (77 % of its peak), where LOP3 / PRMT / VIADDMNMX / VIMNMX issue at half the rate
// of the FMA pipe's IMAD: every LOP3 saved per cell is ~8 % of the kernel.
__device__ __forceinline__ uint32_t sw_opaque(uint32_t v) { uint32_t r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ uint32_t and8_or(uint32_t a, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, 0x00080008, %2, 0xEA;" : "=r"(d) : "r"(a), "r"(c));
    return d;
}
// a * b + c on the FMA pipe (mad.lo): the direction codes (K - m) << 3r are accumulated as K * 2^3r - m * 2^3r
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct StripState {
    uint32_t mleft[SW_R];     // m(i, j-1) of both pairs
    uint32_t left15[SW_R];    // left candidate key for column j: m(i,j-1) + 8*[left] - 15
    uint32_t mx[SW_R];        // running row maxima
};

// columns j_from..j_to of one strip.  CAPPED: a pair whose window ended before this column keeps computing (its lane
// is garbage from here on) but must not raise its row maxima: its keys are clamped to -32768 for the maximum.
// TRACK = false: only the strip's last row keeps its running maximum (see the row-threshold note in sw_fast_kernel).
template <bool MASKED, bool CAPPED, bool DELTA, bool TRACK>
__device__ __forceinline__ void sw_cols(std::conditional_t<DELTA, uint16_t, uint32_t> *__restrict__ row, const uint32_t *__restrict__ refw,
                                        uint2 *__restrict__ dstrip, const int j_from, const int j_to, const int rminA, const int rminB,
                                        const uint32_t capv, const uint32_t (&profA)[SW_R], const uint32_t (&profB)[SW_R],
                                        StripState &S, uint32_t &kt_prev, uint32_t &kl_prev, uint32_t &ovf) {
    constexpr int B = SW_BLOCK;
    if (j_from > j_to) return;
    // lanes whose window has ended (CAPPED) hold garbage: their differences must not raise the overflow flag
    const uint32_t live = CAPPED ? (((capv & 0xFFFFu) == 0x7FFFu ? 0xFFFFu : 0u) | ((capv >> 16) == 0x7FFFu ? 0xFFFF0000u : 0u)) : 0xFFFFFFFFu;
    uint32_t rwA = refw[(2 * ((j_from - 1) >> 4)) * B] >> (2 * ((j_from - 1) & 15));
    uint32_t rwB = refw[(2 * ((j_from - 1) >> 4) + 1) * B] >> (2 * ((j_from - 1) & 15));
    const uint32_t c_up = sw_opaque(0xFFF2FFF2u), c_left = sw_opaque(0xFFF1FFF1u);
    for (int j = j_from; j <= j_to; j++) {
        if (((j - 1) & 15) == 0) { rwA = refw[(2 * ((j - 1) >> 4)) * B]; rwB = refw[(2 * ((j - 1) >> 4) + 1) * B]; }
        // PRMT selector: lane A = byte cA of profA sign-extended, lane B = byte cB of profB sign-extended
        const uint32_t sel = (rwA & 3u) * 0x11u + (rwB & 3u) * 0x1100u + 0xC480u;
        rwA >>= 2; rwB >>= 2;
        // the boundary row is stored as one signed byte per lane: K(i0-1,j) - K(i0-1,j-1)  (|dV| <= 6 => |dK| < 64)
        const uint32_t kt = DELTA ? __vadd2(kt_prev, prmt((uint32_t)row[j * B], 0u, 0x9180u)) : (uint32_t)row[j * B];   // K(i0-1, j)
        const uint32_t mT = kt & 0xFFF8FFF8u;
        const uint32_t mD_top = kt_prev & 0xFFF8FFF8u;                 // m(i0-1, j-1)
        kt_prev = kt;
        uint32_t mup = mT;
        uint32_t tup = and8_or(kt * 4u, c_up);                         // 8*[up] - 14 per lane
        uint32_t mdiag = mD_top;
        uint32_t acc_lo = 0, acc_hi = 0, K = 0;
#pragma unroll
        for (int r = 0; r < SW_R; r++) {
            const uint32_t s = prmt(profA[r], profB[r], sel);          // 8*sigma + 4, both lanes
            const uint32_t x = __viaddmax_s16x2(mdiag, s, S.left15[r]);   // max(diag, left)
            K = __viaddmax_s16x2(mup, tup, x);                          // max(up, .)
            if (MASKED) K &= (r >= rminA ? 0x0000FFFFu : 0u) | (r >= rminB ? 0xFFFF0000u : 0u);
            mdiag = S.mleft[r];
            const uint32_t m = K & 0xFFF8FFF8u;
            S.mleft[r] = m;
            S.left15[r] = __vadd2(m, and8_or(K * 8u, c_left));          // m + 8*[left] - 15
            mup = m;
            tup = and8_or(K * 4u, c_up);
            if (TRACK || r == SW_R - 1) S.mx[r] = __vmaxs2(S.mx[r], CAPPED ? __vmins2(K, capv) : K);
            // the cell's 3-bit direction code K - m at bits 3r of the strip's word
            if (r < 5) { acc_lo = mad_lo(K, 1u << (3 * r), acc_lo); acc_lo = mad_lo(m, 0u - (1u << (3 * r)), acc_lo); }
            else { acc_hi = mad_lo(K, 1u << (3 * (r - 5)), acc_hi); acc_hi = mad_lo(m, 0u - (1u << (3 * (r - 5))), acc_hi); }
        }
        if (DELTA) {
            const uint32_t dk = __vsub2(K, kl_prev);
            const uint32_t d8 = prmt(dk, 0u, 0x0020u);                  // low bytes of both lanes' differences
            ovf |= (prmt(d8, 0u, 0x9180u) ^ dk) & live;                 // a difference that does not fit a signed byte
            row[j * B] = (uint16_t)d8;
            kl_prev = K;
        } else {
            row[j * B] = K;
        }
        dstrip[j * B] = make_uint2(acc_lo, acc_hi);
    }
}

// The two pairs of a thread may have different query lengths: both are aligned at their LAST row, so lane X's row r of
// this strip is query row i0X + r, and rows <= 0 (only in the first strips, more of them for the shorter query) act as
// row 0 of the table.
template <bool MASKED, bool DELTA, bool TRACK>
__device__ __forceinline__ void sw_strip(std::conditional_t<DELTA, uint16_t, uint32_t> *__restrict__ row, const uint32_t *__restrict__ refw,
                                         uint2 *__restrict__ dstrip, const int lr_min, const int lr_max, const uint32_t capv,
                                         const int i0A, const int i0B, const uint32_t (&profA)[SW_R], const uint32_t (&profB)[SW_R],
                                         StripState &S, uint32_t &ovf) {
    const int rminA = 1 - i0A, rminB = 1 - i0B;     // first real row of a masked strip (rows i <= 0 are row 0)
    auto col0 = [](int i) { return i >= 1 ? -8 * (i + 1) + 2 : 0; };      // column 0: V = -(i+1), dir = up   (sw.erl:72-77)
    auto pack2 = [](int a, int b) { return ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16); };
#pragma unroll
    for (int r = 0; r < SW_R; r++) {
        const int kA = col0(i0A + r), kB = col0(i0B + r);
        S.mleft[r] = pack2(kA & ~7, kB & ~7);
        S.left15[r] = __vadd2(S.mleft[r], 0xFFF1FFF1u);
        S.mx[r] = pack2(kA, kB);
    }
    uint32_t kt_prev = pack2(col0(i0A - 1), col0(i0B - 1));              // K(i0-1, 0): column 0 is known in closed form
    uint32_t kl_prev = pack2(col0(i0A + SW_R - 1), col0(i0B + SW_R - 1));    // K(last row of the strip, 0)
    sw_cols<MASKED, false, DELTA, TRACK>(row, refw, dstrip, 1, lr_min, rminA, rminB, 0u, profA, profB, S, kt_prev, kl_prev, ovf);
    sw_cols<MASKED, true, DELTA, TRACK>(row, refw, dstrip, lr_min + 1, lr_max, rminA, rminB, capv, profA, profB, S, kt_prev, kl_prev, ovf);
}

// traceback of one lane (sw.erl:116-160) + CIGAR text; returns false when the run buffer overflowed.
// The direction words live in HBM/L2: instead of one dependent load per step, the SW_TB_BAND columns to the left of
// the current cell are fetched together (independent loads) into the thread's shared-memory slice once per strip
// (again only if the path leaves the band), and the walk reads them from there.
constexpr int SW_TB_BAND = 16;

__device__ __forceinline__ bool sw_traceback(const SwFastParams &p, const uint2 *dirs, uint16_t *scratch, int lane, int lq,
                                             int lr, int i_first, int j, int vmax, uint32_t cand) {
    constexpr int B = SW_BLOCK;
    uint16_t *runs = scratch;                                          // runs[k*B], k < run_cap
    const int run_cap = p.cap_cols - 4 * SW_TB_BAND;                   // the band sits behind the runs
    uint16_t *band = scratch + (size_t)run_cap * B;                    // band[(4c+h)*B], h = 0..3: the uint2 of column band_hi - c
    int i = lq, nr = 0, cur = -1, len = 0;
    bool ovf = false;
    int band_s = -1, band_hi = -1;                                     // strip and rightmost column held in the band
    while (i > 0) {
        if (j < 0) return false;                              // cannot happen with consistent inputs: never spin
        int op;                                               // 0 = M (up_left), 1 = D (left), 2 = I (up)
        if (j == 0) op = 2;
        else {
            const int t = i - i_first, s = t / SW_R, r = t - s * SW_R;
            if (s != band_s || j > band_hi || j <= band_hi - SW_TB_BAND) {
                band_s = s; band_hi = j;
#pragma unroll
                for (int c = 0; c < SW_TB_BAND; c++) {
                    if (j - c >= 1) {
                        const uint2 d = dirs[((size_t)s * p.cap_cols + (j - c)) * B];
                        band[(4 * c) * B] = (uint16_t)d.x; band[(4 * c + 1) * B] = (uint16_t)(d.x >> 16);
                        band[(4 * c + 2) * B] = (uint16_t)d.y; band[(4 * c + 3) * B] = (uint16_t)(d.y >> 16);
                    }
                }
            }
            // rows 0-4 of the strip sit in d.x, rows 5-9 in d.y; each half-word belongs to one lane
            const uint32_t hw = band[(4 * (band_hi - j) + (r < 5 ? 0 : 2) + lane) * B];
            uint32_t code = (hw >> (3 * (r < 5 ? r : r - 5))) & 7u;
            op = code == 4u ? 0 : code == 2u ? 2 : 1;
        }
        if (op == cur) len++;
        else {
            if (cur >= 0) { if (nr < run_cap) runs[nr * B] = (uint16_t)((len << 2) | cur); else ovf = true; nr++; }
            cur = op; len = 1;
        }
        if (op == 0) { i--; j--; } else if (op == 2) i--; else j--;
    }
    if (nr < run_cap) runs[nr * B] = (uint16_t)((len << 2) | cur); else ovf = true;
    nr++;
    if (ovf) return false;
    int clen = 0;
    for (int k = 0; k < nr; k++) {
        int l = (int)(runs[k * B] >> 2);
        clen += 2 + (l >= 10) + (l >= 100) + (l >= 1000) + (l >= 10000);
    }
    bool room;
    uint32_t off = sw_cig_alloc(p.ctr, p.cig_cap, (uint32_t)clen, room);
    if (room) {
        char *c = p.cig + off;
        for (int k = nr - 1; k >= 0; k--) {
            int v = (int)runs[k * B];
            int l = v >> 2, op = v & 3;
            int nd = 1 + (l >= 10) + (l >= 100) + (l >= 1000) + (l >= 10000);
            for (int t = nd - 1; t >= 0; t--) { c[t] = (char)('0' + l % 10); l /= 10; }
            c[nd] = op == 0 ? 'M' : op == 1 ? 'D' : 'I';
            c += nd + 1;
        }
    }
    SwOut o;
    o.score = vmax; o.status = BWT_HIT_MATCH; o.cigar_len = clen; o.cigar_off = off;
    p.out[cand] = o;
    return true;
}

constexpr int SW_MAXQW = 26;    // query code words per lane: ceil(400 / 16) + 1

// stage one pair (both strings plain ACGT: the classification checked): window codes into slot `lane` of refw (shared),
// query codes into qc (registers / local)
__device__ __forceinline__ void sw_stage(const SwFastParams &p, const SwDesc &d, uint32_t *refw, int lane, uint32_t *qc) {
    constexpr int B = SW_BLOCK;
    for (int j0 = 0; j0 < d.lr; j0 += 16) refw[((j0 >> 4) * 2 + lane) * B] = ref_word(p.ref2, d.r_off + j0);
    const uint32_t *pkr = p.R.pk + __ldg(p.R.woff + d.read);
    for (int w = 0; 16 * w < d.lq; w++) qc[w] = read_fwd_word(pkr, d.lq, w);
}

// DELTA: the boundary row is kept as signed-byte differences (half the shared memory, twice the resident warps);
// pairs whose differences overflow are listed in list_ovf and redone by the DELTA = false instantiation.
template <bool DELTA>
__global__ void __launch_bounds__(SW_BLOCK) sw_fast_kernel(SwFastParams p) {
    extern __shared__ uint32_t smem[];
    using RowT = std::conditional_t<DELTA, uint16_t, uint32_t>;
    constexpr int B = SW_BLOCK;
    constexpr int R = SW_R;
    const int tid = threadIdx.x;
    RowT *row = reinterpret_cast<RowT *>(smem) + tid;                                   // row[j*B]: both pairs' keys (or differences)
    uint32_t *refw = smem + (p.cap_cols * B) / (DELTA ? 2 : 1) + tid;                   // refw[(2w+lane)*B]
    uint16_t *scratch = reinterpret_cast<uint16_t *>(smem) + tid;                       // the traceback reuses the row as half-words
    uint2 *dirs = p.dirs + (size_t)blockIdx.x * p.max_strips * p.cap_cols * B + tid;    // dirs[(s*cap+j)*B]
    const uint32_t n_list = *p.n_list;
    const uint32_t n_units = (n_list + 1) / 2;
    unsigned long long calls = 0, cells = 0;
    auto pack2 = [](int a, int b) { return ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16); };

    for (uint32_t unit = blockIdx.x * B + tid; unit < n_units; unit += gridDim.x * B) {
        // two pairs per thread, one in each half of the packed registers; their query lengths may differ (they are
        // aligned at their last rows); a lone last pair is shadowed by itself
        uint32_t candA = p.list[2 * unit];
        uint32_t candB = 2 * unit + 1 < n_list ? p.list[2 * unit + 1] : 0xFFFFFFFFu;
        SwDesc dA = p.desc[candA], dB = candB != 0xFFFFFFFFu ? p.desc[candB] : dA;
        uint32_t qcA[SW_MAXQW], qcB[SW_MAXQW];
        KD_CHECK(dA.lq >= 1 && dA.lr >= 1 && dA.lr + 1 <= p.cap_cols && (dA.lq + R - 1) / R <= p.max_strips, "cand=%u lq=%d lr=%d", candA, dA.lq, dA.lr);
        sw_stage(p, dA, refw, 0, qcA);
        sw_stage(p, dB, refw, 1, qcB);
        const int lqA = dA.lq, lqB = dB.lq, lrA = dA.lr, lrB = dB.lr;
        const int lr_min = min(lrA, lrB), lr_max = max(lrA, lrB);
        const uint32_t capv = (lrA > lr_min ? 0x7FFFu : 0x8000u) | ((lrB > lr_min ? 0x7FFFu : 0x8000u) << 16);
        for (int j = 0; j <= lr_max; j++) row[j * B] = 0;            // row 0: {0,undef} everywhere (sw.erl:61): all differences 0
        const int nstr = (max(lqA, lqB) + R - 1) / R;
        const int ifA = lqA - nstr * R + 1, ifB = lqB - nstr * R + 1;    // row number of strip 0, r = 0 (<= 1), per lane
        const int thrA = -2 * (lqA / 2), thrB = -2 * (lqB / 2);          // sw.erl:62
        int rowsA = lqA, rowsB = lqB;
        bool failA = false, failB = false;
        StripState S;
        int slackA = 0, slackB = 0;                                  // Vmax - threshold at the last row of the previous strip
        uint32_t ovf = 0;                                            // per lane: a boundary-row difference left int8
        for (int s = 0; s < nstr; s++) {
            const int i0A = ifA + s * R, i0B = ifB + s * R;
            uint32_t profA[R], profB[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int iA = i0A + r, iB = i0B + r;
                const uint32_t qa = (iA >= 1) ? ((qcA[(iA - 1) >> 4] >> (2 * ((iA - 1) & 15))) & 3u) : 0u;
                const uint32_t qb = (iB >= 1) ? ((qcB[(iB - 1) >> 4] >> (2 * ((iB - 1) & 15))) & 3u) : 0u;
                profA[r] = 0xFCFCFCFCu ^ (0xE8u << (8 * qa));        // byte qa = 20 (match), others = -4 (mismatch)
                profB[r] = 0xFCFCFCFCu ^ (0xE8u << (8 * qb));
            }
            uint2 *dstrip = dirs + (size_t)s * p.cap_cols * B;
            // Row threshold (sw.erl:69-71): Vmax of row i must reach thr0 + 2*i.  A row's maximum is at most 2 below the
            // maximum of the row above (the `up` candidate of the cell under the best cell costs -1 or -2, sw.erl:105-106)
            // and the threshold rises by 2 per row, so the slack Vmax(i) - thr(i) falls by at most 4 per row: when both
            // pairs left the previous strip with a slack of 4*R or more, no row of this strip can fail, and the ten
            // VIMNMX per column that keep the rows' maxima shrink to the one of the strip's last row.
            const bool light = i0A >= 1 && i0B >= 1 && slackA >= 4 * R && slackB >= 4 * R;
            if (i0A < 1 || i0B < 1) sw_strip<true, DELTA, true>(row, refw, dstrip, lr_min, lr_max, capv, i0A, i0B, profA, profB, S, ovf);
            else if (light) sw_strip<false, DELTA, false>(row, refw, dstrip, lr_min, lr_max, capv, i0A, i0B, profA, profB, S, ovf);
            else sw_strip<false, DELTA, true>(row, refw, dstrip, lr_min, lr_max, capv, i0A, i0B, profA, profB, S, ovf);
            if (!light) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int iA = i0A + r, iB = i0B + r;
                    if (iA >= 1 && !failA && (lane16(S.mx[r], 0) >> 3) < thrA + 2 * iA) { failA = true; rowsA = iA; }
                    if (iB >= 1 && !failB && (lane16(S.mx[r], 1) >> 3) < thrB + 2 * iB) { failB = true; rowsB = iB; }
                }
            }
            if (failA && failB) break;
            // a failed pair keeps the other one going at full tracking cost or not at all: its own slack no longer matters
            slackA = failA ? 0x7FFF : (lane16(S.mx[R - 1], 0) >> 3) - (thrA + 2 * (i0A + R - 1));
            slackB = failB ? 0x7FFF : (lane16(S.mx[R - 1], 1) >> 3) - (thrB + 2 * (i0B + R - 1));
        }
        // a pair whose boundary row could not be difference-coded is redone with absolute rows (rare: it takes
        // a gap run of a dozen bases right next to a high-scoring path)
        bool liveA = true, liveB = candB != 0xFFFFFFFFu;
        if (ovf & 0xFFFFu) { p.list_ovf[atomicAdd(&p.ctr->n_ovf, 1u)] = candA; liveA = false; }
        if ((ovf >> 16) && liveB) { p.list_ovf[atomicAdd(&p.ctr->n_ovf, 1u)] = candB; liveB = false; }
        SwOut o;
        o.score = 0; o.status = BWT_HIT_NO_MATCH; o.cigar_len = 0; o.cigar_off = 0;
        // leftmost maximum of the last row (sw.erl:110-112), for both lanes before the row buffer is reused
        const int vmaxA = lane16(S.mx[R - 1], 0) >> 3, vmaxB = lane16(S.mx[R - 1], 1) >> 3;
        int jA = -1, jB = -1;
        {
            uint32_t k = pack2(-8 * (lqA + 1) + 2, -8 * (lqB + 1) + 2);   // K(lq, 0); then the running sum of the differences
            const bool needA = liveA && !failA, needB = liveB && !failB;
            for (int j = 0; j <= lr_max && ((needA && jA < 0) || (needB && jB < 0)); j++) {
                if (j) k = DELTA ? __vadd2(k, prmt((uint32_t)row[j * B], 0u, 0x9180u)) : (uint32_t)row[j * B];
                if (jA < 0 && j <= lrA && (lane16(k, 0) >> 3) == vmaxA) jA = j;
                if (jB < 0 && j <= lrB && (lane16(k, 1) >> 3) == vmaxB) jB = j;
            }
            KD_CHECK((!needA || jA >= 0) && (!needB || jB >= 0), "cand=%u/%u vmax=%d/%d", candA, candB, vmaxA, vmaxB);
            if (needA && jA < 0) { p.list_gen[atomicAdd(&p.ctr->n_general, 1u)] = candA; liveA = false; }
            if (needB && jB < 0) { p.list_gen[atomicAdd(&p.ctr->n_general, 1u)] = candB; liveB = false; }
        }
        if (liveA) {
            calls++; cells += (unsigned long long)rowsA * (unsigned)lrA;
            if (failA) p.out[candA] = o;
            else if (!sw_traceback(p, dirs, scratch, 0, lqA, lrA, ifA, jA, vmaxA, candA)) {
                p.list_gen[atomicAdd(&p.ctr->n_general, 1u)] = candA; calls--; cells -= (unsigned long long)rowsA * (unsigned)lrA;
            }
        }
        if (liveB) {
            calls++; cells += (unsigned long long)rowsB * (unsigned)lrB;
            if (failB) p.out[candB] = o;
            else if (!sw_traceback(p, dirs, scratch, 1, lqB, lrB, ifB, jB, vmaxB, candB)) {
                p.list_gen[atomicAdd(&p.ctr->n_general, 1u)] = candB; calls--; cells -= (unsigned long long)rowsB * (unsigned)lrB;
            }
        }
    }
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
        calls += __shfl_down_sync(0xFFFFFFFFu, calls, dlt);
        cells += __shfl_down_sync(0xFFFFFFFFu, cells, dlt);
    }
    if ((tid & 31) == 0 && calls) { atomicAdd(&p.ctr->sw_calls, calls); atomicAdd(&p.ctr->cells, cells); }
}

// ---------------------------------------------------------------------------------------------------
// The DP list ordered by shape.  sw_fast_kernel runs two pairs per thread and 32 threads per warp in lock step: a warp
// pays for its longest window (columns), its largest strip count (Lq = 101 needs one strip more than Lq = 100) and for
// BOTH instantiations of the first strip when some of its lanes start on a full strip and others on a masked one.
// With ragged read lengths (98-102 bp) that cost 15 % of the lanes (27 of 32 active in the inner loop).  A counting
// sort by (Lq, Lr) -- histogram, one-block scan, scatter; warp-aggregated atomics because a batch has a few dozen
// shapes -- puts equal shapes next to each other.  The order inside a shape is arbitrary; results do not depend on it.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t SW_SORT_LR = 512;          // window lengths per query length (the fast tiers end at 384 columns)
__device__ __forceinline__ uint32_t sw_shape_key(const SwDesc &d, uint32_t nbins) {
    const uint32_t k = (uint32_t)d.lq * SW_SORT_LR + min((uint32_t)d.lr, SW_SORT_LR - 1u);
    return nbins - 1u - min(k, nbins - 1u);        // largest tables first: the kernel's last, partly filled wave is then made of the cheapest pairs
}

__global__ void __launch_bounds__(256) sw_shape_hist_kernel(const SwDesc *__restrict__ desc, const uint32_t *__restrict__ list,
                                                            const uint32_t *__restrict__ n_ptr, uint32_t *__restrict__ bins, uint32_t nbins) {
    const uint32_t n = *n_ptr;
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t e0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; e0 < n; e0 += gridDim.x * blockDim.x) {
        const uint32_t e = e0 + lane;
        const uint32_t key = e < n ? sw_shape_key(desc[list[e]], nbins) : 0xFFFFFFFFu;
        const unsigned m = __match_any_sync(0xFFFFFFFFu, key);
        if (e < n && lane == (unsigned)(__ffs((int)m) - 1)) atomicAdd(&bins[key], (uint32_t)__popc(m));
    }
}

__global__ void __launch_bounds__(1024) sw_shape_scan_kernel(uint32_t *__restrict__ bins, uint32_t nbins) {
    __shared__ uint32_t s_part[1024];
    const uint32_t tid = threadIdx.x, per = (nbins + 1023u) / 1024u;
    const uint32_t lo = min(nbins, tid * per), hi = min(nbins, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += bins[i];
    s_part[tid] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {
        const uint32_t v = tid >= o ? s_part[tid - o] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (uint32_t i = lo; i < hi; i++) { const uint32_t c = bins[i]; bins[i] = run; run += c; }
}

__global__ void __launch_bounds__(256) sw_shape_scatter_kernel(const SwDesc *__restrict__ desc, const uint32_t *__restrict__ list,
                                                               const uint32_t *__restrict__ n_ptr, uint32_t *__restrict__ bins, uint32_t nbins,
                                                               uint32_t *__restrict__ sorted) {
    const uint32_t n = *n_ptr;
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t e0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; e0 < n; e0 += gridDim.x * blockDim.x) {
        const uint32_t e = e0 + lane;
        const uint32_t k = e < n ? list[e] : 0u;
        const uint32_t key = e < n ? sw_shape_key(desc[k], nbins) : 0xFFFFFFFFu;
        const unsigned m = __match_any_sync(0xFFFFFFFFu, key);
        const int leader = __ffs((int)m) - 1;
        uint32_t base = 0;
        if (e < n && (int)lane == leader) base = atomicAdd(&bins[key], (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (e < n) sorted[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = k;
    }
}

// ---------------------------------------------------------------------------------------------------
// general path: doubles, any alphabet, any size
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sigma_ref(uint8_t q, uint8_t r) {      // sw.erl:164-202, first clause wins
    if (q == 'N' && r == 'N') return 1.6;
    if (q == r) return 2.0;
    if (q == 'N' || r == 'N') return 1.6;
    switch (r) {
    case 'R': return (q == 'A' || q == 'G') ? 2.0 : -1.0;
    case 'Y': return (q == 'C' || q == 'T') ? 2.0 : -1.0;
    case 'B': return (q == 'A') ? -1.0 : 2.0;
    case 'D': return (q == 'C') ? -1.0 : 2.0;
    case 'K': return (q == 'G' || q == 'T') ? 2.0 : -1.0;
    case 'M': return (q == 'A' || q == 'C') ? 2.0 : -1.0;
    case 'S': return (q == 'G' || q == 'C') ? 2.0 : -1.0;
    case 'W': return (q == 'A' || q == 'T') ? 2.0 : -1.0;
    case 'V': return (q == 'T') ? -1.0 : 2.0;
    default: return -1.0;
    }
}

__global__ void sw_dims_kernel(const SwDesc *__restrict__ desc, const uint32_t *__restrict__ list, uint32_t n,
                               int2 *__restrict__ dims) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { SwDesc d = desc[list[k]]; dims[k] = make_int2(d.lq, d.lr); }
}

__host__ __device__ inline uint64_t sw_general_scratch_bytes(int lq, int lr) {
    uint64_t W = (uint64_t)lr + 1;
    uint64_t b = 2 * W * 8 + ((uint64_t)lq + 1) * W + 4 * ((uint64_t)lq + lr + 2);
    return (b + 15) & ~15ull;
}

__global__ void __launch_bounds__(64) sw_general_kernel(DevReads R, const uint8_t *__restrict__ rbase,
                                                        const SwDesc *__restrict__ desc, const uint32_t *__restrict__ list,
                                                        uint32_t first, uint32_t count, const uint64_t *__restrict__ off,
                                                        uint8_t *scratch, SwOut *out, SwCounters *ctr, char *cig,
                                                        uint32_t cig_cap) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t cand = list[first + t];
    const SwDesc d = desc[cand];
    // the read's text: ASCII when it is resident (flagged reads always are), else out of the packed words
    const uint8_t *w = rbase + d.r_off;
    const int lq = d.lq, lr = d.lr, W = lr + 1;
    const uint8_t *q = nullptr;
    const uint32_t *pkr = nullptr;
    if (R.ascii_full || read_is_exc(R, d.read)) { int l2; q = read_ascii(R, d.read, l2); }
    else pkr = R.pk + R.woff[d.read];
    uint8_t *base = scratch + (off[first + t] - off[first]);
    double *pv = reinterpret_cast<double *>(base), *cv = pv + W;
    uint8_t *dirs = reinterpret_cast<uint8_t *>(cv + W);                 // (lq+1) x W, codes 0 left 1 up 2 up_left 3 undef
    uint32_t *runs = reinterpret_cast<uint32_t *>(base + ((2 * (uint64_t)W * 8 + ((uint64_t)lq + 1) * W + 3) & ~3ull));
    for (int j = 0; j < W; j++) { pv[j] = 0.0; dirs[j] = 3; }
    double thr = (double)(-2 * (lq / 2));
    double vmax = 0.0;
    bool failed = false;
    int rows_done = lq;
    for (int i = 1; i <= lq; i++) {
        if (vmax < thr) { failed = true; rows_done = i - 1; break; }      // sw.erl:69-71
        uint8_t *cd = dirs + (size_t)i * W, *pd = cd - W;
        cv[0] = pv[0] + (i == 1 ? -2.0 : -1.0);                           // sw.erl:72-77
        cd[0] = 1;
        double rowmax = cv[0];
        const uint8_t qc = q ? q[i - 1] : (uint8_t)"ACGT"[read_code(pkr, lq, i - 1)];
        for (int j = 1; j < W; j++) {                                     // sw.erl:88-98
            double bv = __dadd_rn(pv[j - 1], sigma_ref(qc, w[j - 1]));
            int bd = 2;
            double c = __dadd_rn(cv[j - 1], cd[j - 1] == 0 ? -1.0 : -2.0);
            if (c > bv) { bv = c; bd = 0; }
            c = __dadd_rn(pv[j], pd[j] == 1 ? -1.0 : -2.0);
            if (c > bv || (c == bv && bd == 0)) { bv = c; bd = 1; }
            cv[j] = bv; cd[j] = (uint8_t)bd;
            if (bv > rowmax) rowmax = bv;
        }
        double *tp = pv; pv = cv; cv = tp;
        vmax = rowmax;
        thr += 2.0;
    }
    if (!failed && vmax < thr) failed = true;
    atomicAdd(&ctr->sw_calls, 1ull);
    atomicAdd(&ctr->cells, (unsigned long long)rows_done * (unsigned)lr);
    atomicAdd(&ctr->general_done, 1u);
    SwOut o;
    o.score = 0; o.status = BWT_HIT_NO_MATCH; o.cigar_len = 0; o.cigar_off = 0;
    if (failed) { out[cand] = o; return; }
    int j = 0;
    while (pv[j] != vmax) j++;
    int i = lq, nr = 0, cur = -1, len = 0;
    while (i > 0) {
        uint8_t dc = dirs[(size_t)i * W + j];
        int op = dc == 2 ? 0 : dc == 1 ? 2 : 1;
        if (op == cur) len++;
        else { if (cur >= 0) runs[nr++] = ((uint32_t)len << 2) | (uint32_t)cur; cur = op; len = 1; }
        if (op == 0) { i--; j--; } else if (op == 2) i--; else j--;
    }
    runs[nr++] = ((uint32_t)len << 2) | (uint32_t)cur;
    int clen = 0;
    for (int k = 0; k < nr; k++) {
        uint32_t l = runs[k] >> 2;
        int nd = 1;
        while (l >= 10) { nd++; l /= 10; }
        clen += nd + 1;
    }
    bool room;
    uint32_t offc = sw_cig_alloc(ctr, cig_cap, (uint32_t)clen, room);
    if (room) {
        char *c = cig + offc;
        for (int k = nr - 1; k >= 0; k--) {
            uint32_t l = runs[k] >> 2;
            int op = runs[k] & 3, nd = 1;
            for (uint32_t x = l; x >= 10; x /= 10) nd++;
            for (int tt = nd - 1; tt >= 0; tt--) { c[tt] = (char)('0' + l % 10); l /= 10; }
            c[nd] = op == 0 ? 'M' : op == 1 ? 'D' : 'I';
            c += nd + 1;
        }
    }
    o.score = (int32_t)(long long)vmax;                                    // trunc/1 (sw.erl:123)
    o.status = BWT_HIT_MATCH; o.cigar_len = clen; o.cigar_off = offc;
    out[cand] = o;
}

template <typename T>
int grow(T *&p, size_t &cap_items, size_t need_items) {
    if (need_items <= cap_items && p) return BWT_OK;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = need_items + need_items / 4 + 16;
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        cap_items = 0;
        bwt_set_error("device allocation of %zu bytes failed", want * sizeof(T));
        return BWT_ERR_NOMEM;
    }
    cap_items = want;
    return BWT_OK;
}

}  // namespace

int sw_stage_reserve(SwStage &st, uint32_t pairs) {
    if (pairs <= st.cap && st.d_desc) return BWT_OK;
    uint32_t want = pairs + pairs / 4 + 1024;
    if (st.d_desc) { cudaFree(st.d_desc); cudaFree(st.d_out); cudaFree(st.d_list_fast); cudaFree(st.d_list_wide); cudaFree(st.d_list_gen); cudaFree(st.d_list_ovf); cudaFree(st.d_list_wfa); cudaFree(st.d_list_big); cudaFree(st.d_list_sorted); cudaFree(st.d_list_gen2); cudaFree(st.d_list_fast2); cudaFree(st.d_list_fast3); }
    st.d_list_sorted = nullptr; st.d_list_gen2 = nullptr; st.d_list_fast2 = nullptr; st.d_list_fast3 = nullptr;
    st.d_desc = nullptr; st.d_out = nullptr; st.d_list_fast = nullptr; st.d_list_wide = nullptr; st.d_list_gen = nullptr; st.d_list_ovf = nullptr; st.d_list_wfa = nullptr; st.d_list_big = nullptr;
    st.cap = 0;
    CU_TRY(cudaMalloc(&st.d_desc, (size_t)want * sizeof(SwDesc)));
    CU_TRY(cudaMalloc(&st.d_out, (size_t)want * sizeof(SwOut)));
    CU_TRY(cudaMalloc(&st.d_list_fast, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_wide, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_gen, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_ovf, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_wfa, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_big, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_sorted, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_gen2, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_fast2, (size_t)want * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&st.d_list_fast3, (size_t)want * sizeof(uint32_t)));
    if (!st.d_ctr) {
        CU_TRY(cudaMalloc(&st.d_ctr, sizeof(SwCounters)));
        CU_TRY(cudaMallocHost(&st.h_ctr, sizeof(SwCounters)));
        CU_TRY(cudaEventCreate(&st.ev_k0));
        CU_TRY(cudaEventCreate(&st.ev_k1));
    }
    st.cap = want;
    if (st.cig_cap < 32u * want) {
        if (st.d_cig_tmp) cudaFree(st.d_cig_tmp);
        st.d_cig_tmp = nullptr;
        st.cig_cap = 32u * want;
        CU_TRY(cudaMalloc(&st.d_cig_tmp, st.cig_cap));
    }
    return BWT_OK;
}

void sw_stage_free(SwStage &st) {
    cudaFree(st.d_desc); cudaFree(st.d_out); cudaFree(st.d_list_fast); cudaFree(st.d_list_wide); cudaFree(st.d_list_gen); cudaFree(st.d_list_ovf);
    cudaFree(st.d_list_wfa); cudaFree(st.d_list_big); cudaFree(st.d_list_sorted); cudaFree(st.d_sort_bins); cudaFree(st.d_list_gen2); cudaFree(st.d_swd_scratch); cudaFree(st.d_gw_scratch); cudaFree(st.d_list_fast2); cudaFree(st.d_list_fast3); cudaFree(st.d_band_scratch);
    cudaFree(st.d_ctr); if (st.h_ctr) cudaFreeHost(st.h_ctr);
    if (st.ev_k0) cudaEventDestroy(st.ev_k0);
    if (st.ev_k1) cudaEventDestroy(st.ev_k1);
    cudaFree(st.d_cig_tmp); cudaFree(st.d_dirs); cudaFree(st.d_gen_scratch); cudaFree(st.d_gen_off); cudaFree(st.d_gen_dims);
    st = SwStage();
}

namespace {
struct SwPlan {        // launch geometry of one configuration (max_lq)
    int cap_narrow, cap_wide, max_strips;
    unsigned grid[3];
};
size_t smem_of(int cap, bool delta) { return (size_t)SW_BLOCK * ((size_t)cap * (delta ? 2 : 4) + (size_t)(cap / 16) * 8); }

int sw_plan(SwStage &st, uint32_t n_cap, int max_lq, SwPlan &pl) {
    // two shared-memory tiers of the fast kernel: "narrow" holds windows up to Lq + 15 columns (almost every
    // candidate: D is the summed gap of the voting chain), "wide" up to Lq + 160; anything wider takes the general path
    const int cap_limit = 384;                           // fast-kernel tier limit: the absolute-row variant needs 4.5 B x columns x 128 threads <= 227 KB
    pl.cap_narrow = std::min(cap_limit, std::max(96, ((max_lq + 1 + 15) + 15) & ~15));   // >= 96: the traceback's band (64 half-words) + runs
    pl.cap_wide = std::min(cap_limit, std::max(96, ((max_lq + 1 + 160) + 15) & ~15));    // D up to 160: the largest slack seen on GL000193.1 is 142
    pl.max_strips = std::max(1, (std::min(max_lq, SW_FAST_MAX_LQ) + SW_R - 1) / SW_R);
    const int caps[3] = {pl.cap_narrow, pl.cap_wide, pl.cap_wide};
    if (st.cfg_max_lq != max_lq) {                       // attribute / occupancy queries once per configuration, not per batch
        CU_TRY(cudaDeviceGetAttribute(&st.cfg_sms, cudaDevAttrMultiProcessorCount, st.device));
        CU_TRY(cudaFuncSetAttribute(sw_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(pl.cap_wide, true)));
        CU_TRY(cudaFuncSetAttribute(sw_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(pl.cap_wide, false)));
        for (int t = 0; t < 3; t++) {
            if (t < 2) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st.cfg_per_sm[t], sw_fast_kernel<true>, SW_BLOCK, smem_of(caps[t], true)));
            else CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st.cfg_per_sm[t], sw_fast_kernel<false>, SW_BLOCK, smem_of(caps[t], false)));
            if (st.cfg_per_sm[t] < 1) { bwt_set_error("sw_fast_kernel does not fit an SM with %zu B of shared memory", smem_of(caps[t], t < 2)); return BWT_ERR_CUDA; }
        }
        st.cfg_max_lq = max_lq;
    }
    size_t need_dirs = 0;
    for (int t = 0; t < 3; t++) {
        const uint64_t units = ((uint64_t)n_cap + 1) / 2;
        pl.grid[t] = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((units + SW_BLOCK - 1) / SW_BLOCK, (uint64_t)st.cfg_sms * st.cfg_per_sm[t]));
        if (t == 2) pl.grid[t] = std::min<unsigned>(pl.grid[t], (unsigned)st.cfg_sms);       // the overflow list is short: one wave
        need_dirs = std::max(need_dirs, (size_t)pl.grid[t] * SW_BLOCK * pl.max_strips * caps[t]);
    }
    return grow(st.d_dirs, st.dirs_words, need_dirs);
}

SwFastParams sw_params(SwStage &st, const DevReads &R, const uint32_t *d_ref2, const SwPlan &pl) {
    SwFastParams p;
    p.R = R; p.ref2 = d_ref2; p.desc = st.d_desc; p.out = st.d_out;
    p.list_gen = st.d_list_gen; p.list_ovf = st.d_list_ovf; p.ctr = st.d_ctr; p.cig = st.d_cig_tmp; p.cig_cap = st.cig_cap;
    p.dirs = st.d_dirs; p.max_strips = pl.max_strips;
    p.list = nullptr; p.n_list = nullptr; p.cap_cols = 0;
    return p;
}

// BWT_SW_SORT: smallest pair capacity for which the DP list is ordered by shape (default 32768; -1 = never, 0 = always)
int sw_sort_min_pairs() {
    const char *e = getenv("BWT_SW_SORT");
    return e && *e ? atoi(e) : 32768;
}

}  // namespace
// BWT_SW_SMALL: largest batch (reads, or pairs of a bwt_sw_batch call) whose DP pairs all take the warp-per-pair kernel
// (default 4096; 0 = never: the fast tiers even for a handful of pairs)
int sw_small_limit() {
    const char *e = getenv("BWT_SW_SMALL");
    return e && *e ? atoi(e) : 4096;
}
namespace {

// BWT_SW_DOUBLE=0: every double-precision pair through the warp-per-pair kernel (A/B, tests of that kernel)
bool sw_double_enabled() {
    const char *e = getenv("BWT_SW_DOUBLE");
    return !(e && *e && atoi(e) == 0);
}

// BWT_SW_BAND=0: no banded table -- every pair the shortcut leaves open fills its whole table (A/B, parity tests of the
// full-table kernels on gapped pairs)
bool sw_band_enabled() {
    const char *e = getenv("BWT_SW_BAND");
    return !(e && *e && atoi(e) == 0);
}

// BWT_SW_BAND_SORT=0: the list is ordered by shape behind the banded table (for sw_fast_kernel) instead of in front of it
bool sw_band_sort_first() {
    const char *e = getenv("BWT_SW_BAND_SORT");
    return !(e && *e && atoi(e) == 0);
}

int sw_use_shortcut() {
    // BWT_SW_SHORTCUT: 3 (default) quick test + wavefront rule, 2 the wavefront rule alone, 1 the quick test alone,
    // 0 every pair through the DP (A/B timing, and parity tests of the DP on easy pairs); read per run: tests flip it
    const char *e = getenv("BWT_SW_SHORTCUT");
    if (!e || !*e) return 3;
    const int v = atoi(e);
    return v < 0 || v > 3 ? 3 : v;
}
}  // namespace

// Everything of the stage whose launch geometry the host knows without looking at the data: classification (quick test,
// wavefront rule) and the narrow DP tier, all sized by n_cap and driven by device-side counts (*d_n pairs).  No host
// synchronisation: the caller copies st.d_ctr to st.h_ctr behind it and calls sw_stage_finish once that has arrived.
int sw_stage_enqueue(SwStage &st, const DevReads &R, const uint8_t *d_rbase, const uint32_t *d_ref2, const uint32_t *d_refx,
                     const uint32_t *d_n, uint32_t n_cap, int max_lq, cudaStream_t stream) {
    (void)d_rbase;
    st.sw_calls = st.cells = st.n_general = st.shortcut_calls = st.shortcut_cells = 0;
    st.band_calls = st.band_cells = st.band_cells_filled = 0;
    st.ms_fast_kernel = 0;
    if (n_cap > st.cap) { bwt_set_error("sw_stage: %u pairs exceed the reserved %u", n_cap, st.cap); return BWT_ERR_ARG; }
    SwPlan pl;
    BWT_TRY(sw_plan(st, n_cap, max_lq, pl));
    const int use_shortcut = sw_use_shortcut();
    // worker-sized batches (st.small): the thread-per-two-pairs DP kernel is a throughput design -- one pair costs it
    // ~0.4 ms however short the list, and its launch alone (100 KB of shared memory per block) a quarter of that -- so
    // every pair that needs the table goes to the warp-per-pair kernel (tens of microseconds for a few hundred pairs;
    // same recurrence in doubles, exact for plain ACGT as well) and the fast tiers are not launched at all.
    const int cap_narrow = st.small ? 0 : pl.cap_narrow, cap_wide = st.small ? 0 : pl.cap_wide;
    CU_TRY(cudaMemsetAsync(st.d_ctr, 0, sizeof(SwCounters), stream));
    const unsigned cgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n_cap + 255) / 256, (uint64_t)148 * 32));
    sw_classify_kernel<<<cgrid, 256, 0, stream>>>(R, d_ref2, d_refx, st.d_desc, st.d_out, d_n, cap_narrow, cap_wide,
                                                  st.d_list_fast, st.d_list_wide, st.d_list_gen, st.d_list_wfa, st.d_ctr,
                                                  st.d_cig_tmp, st.cig_cap, use_shortcut);
    K_CHECK(stream, "sw_classify_kernel");
    st.launches += 1;
    if (use_shortcut & 2) {
        sw_wfa_kernel<<<std::min(cgrid, 148u * 8u), 256, 0, stream>>>(R, d_ref2, st.d_desc, st.d_out, st.d_list_wfa, cap_narrow, cap_wide,
                                                                     st.d_list_fast, st.d_list_wide, st.d_list_gen, st.d_ctr,
                                                                     st.d_cig_tmp, st.cig_cap);
        K_CHECK(stream, "sw_wfa_kernel");
        st.launches += 1;
    }
    SwFastParams p = sw_params(st, R, d_ref2, pl);
    const uint32_t *fast_list = st.d_list_fast, *fast_n = &st.d_ctr->n_fast;
    const bool band = !st.small && sw_band_enabled();
    const bool sort = !st.small && sw_sort_min_pairs() >= 0 && n_cap >= (uint32_t)sw_sort_min_pairs();
    const bool sort_first = sw_band_sort_first();
    // large batches: the list ordered by shape (see sw_shape_hist_kernel); small ones are not worth three launches.
    // In front of the banded table it keeps a warp's pairs at one band width and one height; what the band leaves comes
    // out in nearly that order, so sw_fast_kernel's short list is not sorted again.
    auto sort_list = [&]() -> int {
        const uint32_t nbins = (uint32_t)(std::min(max_lq, SW_FAST_MAX_LQ) + 1) * SW_SORT_LR;
        BWT_TRY(grow(st.d_sort_bins, st.sort_bins_cap, (size_t)nbins));
        CU_TRY(cudaMemsetAsync(st.d_sort_bins, 0, (size_t)nbins * sizeof(uint32_t), stream));
        const unsigned sgrid = std::min(cgrid, 148u * 8u);
        sw_shape_hist_kernel<<<sgrid, 256, 0, stream>>>(st.d_desc, fast_list, fast_n, st.d_sort_bins, nbins);
        sw_shape_scan_kernel<<<1, 1024, 0, stream>>>(st.d_sort_bins, nbins);
        sw_shape_scatter_kernel<<<sgrid, 256, 0, stream>>>(st.d_desc, fast_list, fast_n, st.d_sort_bins, nbins, st.d_list_sorted);
        K_CHECK(stream, "sw_shape_* kernels");
        st.launches += 3;
        fast_list = st.d_list_sorted;
        return BWT_OK;
    };
    if (sort && (!band || sort_first)) BWT_TRY(sort_list());
    // gapped pairs: a band of the table plus a certificate (swb.cu), tried with a small penalty budget (narrow band) first
    // and with a larger one on what that left; what neither certifies goes on to the full table
    if (band) {
        SwbPlan bp;
        BWT_TRY(sw_band_plan(max_lq, st.cfg_sms, n_cap, &bp));
        BWT_TRY(grow(st.d_band_scratch, st.band_scratch_words, (size_t)bp.scratch_words));
        BWT_TRY(launch_sw_band(bp, R, d_ref2, st.d_desc, fast_list, fast_n, st.d_list_fast2, &st.d_ctr->n_fast2, st.d_band_scratch,
                               st.d_out, st.d_ctr, st.d_cig_tmp, st.cig_cap, stream));
        st.launches += 1;
        fast_list = st.d_list_fast2; fast_n = &st.d_ctr->n_fast2;
        if (bp.p_second) {
            bp.p_prev = bp.p_short; bp.p_short = bp.p_second;
            BWT_TRY(launch_sw_band(bp, R, d_ref2, st.d_desc, fast_list, fast_n, st.d_list_fast3, &st.d_ctr->n_fast3, st.d_band_scratch,
                                   st.d_out, st.d_ctr, st.d_cig_tmp, st.cig_cap, stream));
            st.launches += 1;
            fast_list = st.d_list_fast3; fast_n = &st.d_ctr->n_fast3;
        }
        if (sort && !sort_first) BWT_TRY(sort_list());
    }
    CU_TRY(cudaEventRecord(st.ev_k0, stream));
    if (!st.small) {
        p.list = fast_list; p.n_list = fast_n; p.cap_cols = pl.cap_narrow;
        sw_fast_kernel<true><<<pl.grid[0], SW_BLOCK, smem_of(pl.cap_narrow, true), stream>>>(p);
        K_CHECK(stream, "sw_fast_kernel (narrow)");
        st.launches += 1;
    }
    CU_TRY(cudaEventRecord(st.ev_k1, stream));
    // pairs with 'N' / IUPAC / other bytes, or beyond the fast tiers: one warp per pair, doubles; always launched (its
    // blocks leave at once when the list is empty)
    SwGwPlan gw;
    BWT_TRY(sw_general_warp_plan(max_lq, st.cfg_sms, &gw));
    BWT_TRY(grow(st.d_gw_scratch, st.gw_scratch_words, (size_t)gw.scratch_words));
    const uint32_t *gen_list = st.d_list_gen, *gen_n = &st.d_ctr->n_general;
    if (!st.small && sw_double_enabled()) {
        // throughput path of the double-precision pairs: one thread per pair; what it cannot take (windows wider than its
        // slices, bytes outside the IUPAC letters) goes on to the warp-per-pair kernel
        SwdPlan dp;
        BWT_TRY(sw_double_plan(max_lq, st.cfg_sms, n_cap, &dp));
        BWT_TRY(grow(st.d_swd_scratch, st.swd_scratch_bytes, (size_t)dp.scratch_bytes));
        BWT_TRY(launch_sw_double(dp, R, d_rbase, st.d_desc, st.d_list_gen, &st.d_ctr->n_general, st.d_list_gen2, &st.d_ctr->n_gen2,
                                 st.d_swd_scratch, st.d_out, st.d_ctr, st.d_cig_tmp, st.cig_cap, stream));
        st.launches += 1;
        gen_list = st.d_list_gen2; gen_n = &st.d_ctr->n_gen2;
    }
    BWT_TRY(launch_sw_general_warp(gw, R, d_rbase, st.d_desc, gen_list, gen_n, 0, &st.d_ctr->gen_cursor, st.d_list_big, st.d_gw_scratch,
                                   st.d_out, st.d_ctr, st.d_cig_tmp, st.cig_cap, stream));
    st.launches += 1;
    return BWT_OK;
}

// With st.h_ctr = the counters after sw_stage_enqueue: the rare tiers (wide windows, absolute rows, doubles), each
// launched only when its list is not empty.  *redo_emit is set when any of them ran (results changed after the caller's
// optimistic emit).  BWT_ERR_NOSPACE: the temporary CIGAR arena overflowed; it has been grown, run the stage again.
int sw_stage_finish(SwStage &st, const DevReads &R, const uint8_t *d_rbase, const uint32_t *d_ref2, const uint32_t *d_refx,
                    uint32_t n_cap, int max_lq, cudaStream_t stream, bool *redo_emit) {
    (void)d_refx;
    *redo_emit = false;
    SwPlan pl;
    BWT_TRY(sw_plan(st, n_cap, max_lq, pl));
    SwFastParams p = sw_params(st, R, d_ref2, pl);
    auto grid_for = [&](uint32_t pairs, unsigned cap_grid) {
        return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((((uint64_t)pairs + 1) / 2 + SW_BLOCK - 1) / SW_BLOCK, cap_grid));
    };
    const uint32_t done_general = st.h_ctr->n_general;      // what the always-launched general kernel saw
    auto refresh = [&]() -> int {
        CU_TRY(cudaMemcpyAsync(st.h_ctr, st.d_ctr, sizeof(SwCounters), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        return BWT_OK;
    };
    if (!st.h_ctr->cig_overflow && st.h_ctr->n_wide > 0) {
        p.list = st.d_list_wide; p.n_list = &st.d_ctr->n_wide; p.cap_cols = pl.cap_wide;
        sw_fast_kernel<true><<<grid_for(st.h_ctr->n_wide, pl.grid[1]), SW_BLOCK, smem_of(pl.cap_wide, true), stream>>>(p);
        K_CHECK(stream, "sw_fast_kernel (wide)");
        st.launches += 1;
        *redo_emit = true;
        BWT_TRY(refresh());
    }
    if (!st.h_ctr->cig_overflow && st.h_ctr->n_ovf > 0) {
        // pairs whose boundary rows did not fit the difference coding: absolute rows (rare)
        p.list = st.d_list_ovf; p.n_list = &st.d_ctr->n_ovf; p.cap_cols = pl.cap_wide;
        sw_fast_kernel<false><<<grid_for(st.h_ctr->n_ovf, pl.grid[2]), SW_BLOCK, smem_of(pl.cap_wide, false), stream>>>(p);
        K_CHECK(stream, "sw_fast_kernel (absolute rows)");
        st.launches += 1;
        *redo_emit = true;
        BWT_TRY(refresh());
    }
    if (!st.h_ctr->cig_overflow && st.h_ctr->n_general > done_general) {
        // pairs the wide / absolute-row kernels handed on after the general kernel had run
        SwGwPlan gw;
        BWT_TRY(sw_general_warp_plan(max_lq, st.cfg_sms, &gw));
        CU_TRY(cudaMemsetAsync(&st.d_ctr->gen_cursor, 0, sizeof(uint32_t), stream));
        BWT_TRY(launch_sw_general_warp(gw, R, d_rbase, st.d_desc, st.d_list_gen, &st.d_ctr->n_general, done_general, &st.d_ctr->gen_cursor, st.d_list_big,
                                       st.d_gw_scratch, st.d_out, st.d_ctr, st.d_cig_tmp, st.cig_cap, stream));
        st.launches += 1;
        *redo_emit = true;
        BWT_TRY(refresh());
    }
    const uint32_t ng = st.h_ctr->n_big;
    if (ng > 0 && !st.h_ctr->cig_overflow) {
        // pairs too large for the warp kernel's slots: one thread per pair, scratch sized on the host
        if (ng + 1 > st.gen_off_cap) {
            if (st.d_gen_off) cudaFree(st.d_gen_off);
            if (st.d_gen_dims) cudaFree(st.d_gen_dims);
            st.d_gen_off = nullptr; st.d_gen_dims = nullptr;
            const size_t cap_items = (size_t)ng + ng / 2 + 64;
            CU_TRY(cudaMalloc(&st.d_gen_off, cap_items * sizeof(uint64_t)));
            CU_TRY(cudaMalloc(&st.d_gen_dims, cap_items * sizeof(int2)));
            st.gen_off_cap = cap_items;
        }
        sw_dims_kernel<<<(ng + 255) / 256, 256, 0, stream>>>(st.d_desc, st.d_list_big, ng, st.d_gen_dims);
        st.launches += 1;
        std::vector<int2> dims(ng);
        CU_TRY(cudaMemcpyAsync(dims.data(), st.d_gen_dims, ng * sizeof(int2), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        std::vector<uint64_t> off(ng + 1);
        off[0] = 0;
        for (uint32_t k = 0; k < ng; k++) off[k + 1] = off[k] + sw_general_scratch_bytes(dims[k].x, dims[k].y);
        CU_TRY(cudaMemcpyAsync(st.d_gen_off, off.data(), (ng + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
        // rounds of at most ~2 GB of scratch
        const uint64_t round_bytes = 2ull << 30;
        uint32_t first = 0;
        while (first < ng) {
            uint32_t last = first + 1;
            while (last < ng && off[last + 1] - off[first] <= round_bytes) last++;
            uint64_t need = off[last] - off[first];
            BWT_TRY(grow(st.d_gen_scratch, st.gen_scratch_bytes, (size_t)need));
            uint32_t cnt = last - first;
            sw_general_kernel<<<(cnt + 63) / 64, 64, 0, stream>>>(R, d_rbase, st.d_desc, st.d_list_big, first, cnt,
                                                                 st.d_gen_off, st.d_gen_scratch, st.d_out, st.d_ctr,
                                                                 st.d_cig_tmp, st.cig_cap);
            st.launches += 1;
            K_CHECK(stream, "sw_general_kernel");
            CU_TRY(cudaStreamSynchronize(stream));
            first = last;
        }
        *redo_emit = true;
        BWT_TRY(refresh());
    }
    if (!st.h_ctr->cig_overflow) {
        cudaEventElapsedTime(&st.ms_fast_kernel, st.ev_k0, st.ev_k1);
        st.sw_calls = st.h_ctr->sw_calls; st.cells = st.h_ctr->cells; st.n_general = st.h_ctr->n_general;
        st.shortcut_calls = st.h_ctr->shortcut_calls; st.shortcut_cells = st.h_ctr->shortcut_cells;
        st.band_calls = st.h_ctr->band_calls; st.band_cells = st.h_ctr->band_cells; st.band_cells_filled = st.h_ctr->band_cells_filled;
        return BWT_OK;
    }
    // temporary CIGAR arena too small: grow to what this attempt asked for (and then some); the caller redoes the stage
    uint64_t want = (uint64_t)st.cig_cap * 2 + 4096;                     // a region of the arena ran full
    if (want > 0xFFFFFFF0ull) { bwt_set_error("CIGAR arena beyond 4 GiB for one batch: use smaller batches"); return BWT_ERR_LIMIT; }
    cudaFree(st.d_cig_tmp);
    st.d_cig_tmp = nullptr;
    st.cig_cap = (uint32_t)want;
    CU_TRY(cudaMalloc(&st.d_cig_tmp, st.cig_cap));
    return BWT_ERR_NOSPACE;
}

// enqueue + finish with the synchronisations in between (one-shot callers: bwt_sw_batch)
int sw_stage_run(SwStage &st, const DevReads &R, const uint8_t *d_rbase, const uint32_t *d_ref2, const uint32_t *d_refx,
                 const uint32_t *d_n, uint32_t n, int max_lq, cudaStream_t stream) {
    if (n == 0) return BWT_OK;
    st.small = n <= (uint32_t)std::max(0, sw_small_limit());
    for (int attempt = 0; attempt < 4; attempt++) {
        BWT_TRY(sw_stage_enqueue(st, R, d_rbase, d_ref2, d_refx, d_n, n, max_lq, stream));
        CU_TRY(cudaMemcpyAsync(st.h_ctr, st.d_ctr, sizeof(SwCounters), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        bool redo = false;
        const int rc = sw_stage_finish(st, R, d_rbase, d_ref2, d_refx, n, max_lq, stream, &redo);
        if (rc != BWT_ERR_NOSPACE) return rc;
    }
    bwt_set_error("CIGAR arena kept overflowing");
    return BWT_ERR_NOMEM;
}
