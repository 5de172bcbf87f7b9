// Bounded-edit FM search: the position set of the module that seeds:generate_fs/3 writes (seeds.erl:11-62).
//
//   seeds:generate_fs(Qsec, Len, Defects)   seeds.erl:11-47   -> one matches/Len clause per pattern of gen_pattern/5
//   gen_pattern/5                           seeds.erl:51-62   -> patterns over the first Len+Defects query characters
//                                                                with EXACTLY `Defects` edits: `_` replacing a query
//                                                                character (mutation), `_` inserted, character dropped
//   fs:find_seeds/1                         seeds.erl:23-43   -> greedy left-to-right, non-overlapping window walk
//
// The reference scans the chunk text window by window; here the FM index answers "which windows match some
// pattern" directly.  One warp per query walks gen_pattern/5's recursion from the pattern's RIGHT end (an FM index
// extends matches to the left) as a depth-first search whose states live on a per-warp stack in shared memory:
//
//   state (p, q, d, m, last, [lo,hi)) : p pattern elements still to emit, q query characters still to read, d defects
//                                       left, m matches made so far (capped at Defects), rows [lo,hi) of the suffix array
//   match      c = Q[q]                  -> (p-1, q-1, d,   m+1)       LF step with c
//   mutation   c != Q[q]                 -> (p-1, q-1, d-1, m)         LF step with c   (seeds.erl:58)
//   insertion  any c                     -> (p-1, q,   d-1, m)         LF step with c   (seeds.erl:59)
//   deletion   (only once p < Len)       -> (p,   q-1, d-1, m)         same rows        (seeds.erl:60 + guard :57)
//   accept at p == 0  <=>  q <= d  and  d - q <= m
//
// A `_` also fits the query's own character, so a mutation that keeps the character is a match that spent a defect:
// instead of walking those duplicates, acceptance lets up to m leftover defects be "burnt" on matches already made
// (and q leftover characters dropped by deletions in front of the pattern).  A deletion next to an insertion equals
// a mutation/match that spends fewer defects; once m == Defects that cheaper path is certain to accept whenever the
// dearer one does, so the pair is skipped.  Both rules only remove paths whose windows another path still reaches:
// the window set is exactly the clause list's (tests/test_gpu_inexact.py holds it against a literal CPU restatement).
//
// Each loop iteration pops as many states as fill the warp (10 lanes for a state with defects left: match, 4
// mutations, 4 insertions, deletion; 1 lane for a state without), every lane does its child's LF step (two 32-byte
// Occ sectors), surviving children are pushed with a ballot compaction.  The stack's bottom half spills to HBM when
// shared memory runs out.  Accepted rows are located through the full SA, sorted, made unique and walked greedily by
// the same warp (fs:find_seeds/1), then written in the reference's descending order.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "search.cuh"

namespace {

constexpr int FS_WARPS = 8;              // warps per CTA
constexpr int FS_STACK = 128;            // states per warp in shared memory (16 B each)
constexpr int FS_SPILL = 4096;           // states per warp in the HBM spill area
constexpr int FS_MAX_D = 8;
constexpr int FS_MAX_Q = 1024 + FS_MAX_D; // Len + Defects characters staged per warp

enum { OP_NONE = 0, OP_INS = 1, OP_DEL = 2 };

struct FsCounters {
    unsigned long long nodes, lf_steps, raw_hits, seeds, spills, settled;
    unsigned int next_query;
    unsigned int pad;
};

struct FsShared {
    uint4 stack[FS_WARPS][FS_STACK];
    int8_t q[FS_WARPS][FS_MAX_Q];        // 2-bit code of each staged query character, -1 = not ACGT
};

__device__ __forceinline__ uint32_t pack_pq(int p, int q) { return (uint32_t)p | ((uint32_t)q << 16); }
__device__ __forceinline__ uint32_t pack_dm(int d, int m, int last) { return (uint32_t)d | ((uint32_t)m << 8) | ((uint32_t)last << 16); }

// rows [lo,hi) (0-based, half open) of the suffixes that start with X  ->  rows of the suffixes that start with cX
__device__ __forceinline__ void lf_extend(const DevIndex &ix, int c, uint32_t lo, uint32_t hi, uint32_t &nlo, uint32_t &nhi) {
    const uint32_t base = (c == 0 ? ix.P[0] : c == 1 ? ix.P[1] : c == 2 ? ix.P[2] : ix.P[3]) - 1;   // rows whose first character is smaller than c
    nlo = base + (lo ? occ_at(ix, lo, c).occ : 0u);
    nhi = base + occ_at(ix, hi, c).occ;
}

// ascending bitonic sort of v[0..n2), n2 a power of two, by one warp; v is global memory private to the warp
__device__ void warp_bitonic(uint32_t *v, int n2, int lane) {
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n2; i += 32) {
                int l = i ^ j;
                if (l > i) {
                    uint32_t a = v[i], b = v[l];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { v[i] = b; v[l] = a; }
                }
            }
            __syncwarp();
        }
}

// sort v[0..raw) ascending and drop duplicates, by one warp; v holds at least the next power of two >= raw entries
__device__ uint32_t warp_sort_unique(uint32_t *v, uint32_t raw, int lane) {
    if (raw <= 1) return raw;
    int n2 = 2;
    while ((uint32_t)n2 < raw) n2 <<= 1;
    for (uint32_t k = raw + lane; k < (uint32_t)n2; k += 32) v[k] = 0xFFFFFFFFu;
    __syncwarp();
    warp_bitonic(v, n2, lane);
    uint32_t kept = 0;
    for (uint32_t base = 0; base < raw; base += 32) {
        const uint32_t i = base + lane;
        uint32_t x = 0;
        bool keep = false;
        if (i < raw) { x = v[i]; keep = i == 0 || v[i - 1] != x; }
        __syncwarp();
        const unsigned km = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) v[kept + __popc(km & ((1u << lane) - 1u))] = x;
        kept += __popc(km);
        __syncwarp();
    }
    return kept;
}

// ---------------------------------------------------------------------------------------------------
// Pigeonhole front end (Defects 1 or 2).  A pattern of gen_pattern/5 carries exactly Defects edits -- `_` elements and
// dropped query characters -- so of Defects+1 equal pieces of the pattern at least one holds neither a `_` nor a drop
// between two of its elements: that piece of a matching window is a verbatim copy of the query, Q[a+s, b+s) for the
// piece [a, b) of the pattern and a shift s = drops - insertions in front of it, |s| <= Defects.  So instead of the
// depth-first search, whose first ~14 steps branch into every 2-edit neighbour that still occurs in the text:
//   * (Defects+1)(2 Defects+1) <= 15 EXACT backward searches, one per lane;
//   * every occurrence names one candidate window, start = SA - a;
//   * each distinct candidate is verified against the text by a bit-parallel run of the SAME state machine the
//     depth-first search walks (states (diff = q-p, d, m) in one 64-bit word per pattern position).
// The verified set is exactly the clause list's window set (complete by the pigeonhole argument, sound by the
// verification).  Anything the front end does not cover falls back to the depth-first search of the same query: a
// piece that occurs more than FS_PIG_WIDTH times (repeats), a candidate that touches a byte of the .ref outside ACGT
// (the index holds a substituted letter there, fm_index.erl:131-137, which only the FM walk sees), more candidates
// than the caller's scratch.  BWT_FS_PIGEON=0 switches it off (tests run both ways).
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t FS_PIG_WIDTH = 32;

// One candidate window W = text[s0, s0+len) against the patterns of (Q, len, D), D = 1 or 2: the transitions of
// fs_search_kernel with the text character known instead of searched for, all states of one pattern position in ONE
// 64-bit word.  State (diff = q - p + D in 0..2D, d defects left, m matches made, capped at D) sits at bit
// (diff * (D+1) + d) * (D+1) + m, so every transition is a mask and a shift:
//   match      (q >= 1, Q[q-1] = t)    m -> min(m+1, D)           bits with m < D move up one place
//   `_` for the query character (Q[q-1] != t), d -> d-1           >> (D+1)
//   `_` inserted: p-1 with the same q, diff+1, d-1                << (D+1)^2 - (D+1)
//   drop (p < Len, q >= 1): diff-1, d-1                           >> (D+1)^2 + (D+1), repeated D times
// accept at p = 0:  q <= d  and  d - q <= m.
template <int D>
__device__ bool fs_verify_window_t(const uint8_t *__restrict__ W, const uint8_t *__restrict__ Q, const int len) {
    constexpr int DD = D + 1, NDI = 2 * D + 1, GRP = DD * DD;          // bits per diff group
    constexpr uint64_t ONE = 1ull;
    // masks over the whole word
    uint64_t M_LT = 0, M_DGE1 = 0, M_DIGE1 = 0, M_DILT = 0, M_ACC = 0;
#pragma unroll
    for (int di = 0; di < NDI; di++)
#pragma unroll
        for (int d = 0; d <= D; d++)
#pragma unroll
            for (int m = 0; m <= D; m++) {
                const uint64_t b = ONE << ((di * DD + d) * DD + m);
                if (m < D) M_LT |= b;
                if (d >= 1) M_DGE1 |= b;
                if (di >= 1) M_DIGE1 |= b;
                if (di + 1 < NDI) M_DILT |= b;
                const int q = di - D;
                if (q >= 0 && q <= d && d - q <= m) M_ACC |= b;
            }
    constexpr uint64_t GMASK = (ONE << GRP) - 1;
    uint64_t S = 0;
#pragma unroll
    for (int sft = 0; sft < NDI; sft++)
        if (len - D + sft >= 0) S |= ONE << ((sft * DD + D) * DD);    // pattern ends after query character Len - D + sft; d = D, m = 0
    for (int p = len; p >= 1; p--) {
        // per diff group: does query character q = p + di - D exist, and does it equal the text character?
        const uint8_t t = W[p - 1];
        uint64_t HASQ = 0, EQ = 0;
#pragma unroll
        for (int di = 0; di < NDI; di++) {
            const int q = p + di - D;
            if (q >= 1) {
                HASQ |= GMASK << (di * GRP);
                const uint8_t qch = Q[q - 1];
                if (qch == t && bwt_code(qch) >= 0) EQ |= GMASK << (di * GRP);
            }
        }
        if (p < len) {
#pragma unroll
            for (int rep = 0; rep < D; rep++) S |= (S & M_DGE1 & M_DIGE1 & HASQ) >> (GRP + DD);
        }
        const uint64_t Sm = S & EQ;
        uint64_t N = ((Sm & M_LT) << 1) | (Sm & ~M_LT);                 // match
        N |= (S & HASQ & ~EQ & M_DGE1) >> DD;                           // `_` in place of the query character
        N |= (S & M_DGE1 & M_DILT) << (GRP - DD);                       // `_` inserted
        S = N;
        if (!S) return false;
    }
    return (S & M_ACC) != 0;
}

__device__ __forceinline__ bool fs_verify_window(const uint8_t *__restrict__ W, const uint8_t *__restrict__ Q, const int len, const int D) {
    return D == 1 ? fs_verify_window_t<1>(W, Q, len) : fs_verify_window_t<2>(W, Q, len);
}

// -> true: my_hits[0, *raw_out) holds the matching window starts (1-based, ascending, distinct).  false: run the search.
__device__ bool fs_pigeonhole(const DevIndex &ix, const uint8_t *__restrict__ Q, const int8_t *qc, const int len, const int D,
                              uint32_t *my_hits, const uint32_t cap, const int lane, uint32_t &raw_out,
                              unsigned long long &c_lf, unsigned long long &c_nodes) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int NDI = 2 * D + 1, nsearch = (D + 1) * NDI;
    const uint32_t n_rows = ix.n;
    // ---- exact search of one piece at one shift per lane ----
    uint32_t lo = 0, hi = 0;
    int a = 0;
    if (lane < nsearch) {
        const int j = lane / NDI, sft = lane % NDI - D;
        a = (int)((long long)j * len / (D + 1));
        const int b = (int)((long long)(j + 1) * len / (D + 1));
        const int qa = a + sft, qb = b + sft;
        if (qa >= 0 && qb <= len + D && b > a) {
            lo = 0; hi = n_rows;
            for (int x = qb - 1; x >= qa && lo < hi; x--) {
                const int c = (int)qc[x];
                if (c < 0) { hi = lo; break; }
                uint32_t nlo, nhi;
                lf_extend(ix, c, lo, hi, nlo, nhi);
                lo = nlo; hi = nhi;
                c_lf++;
            }
        }
    }
    const uint32_t width = hi > lo ? hi - lo : 0u;
    if (__any_sync(FULL, width > FS_PIG_WIDTH)) return false;
    uint32_t incl = width;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    if (total == 0) { c_nodes += (lane == 0) ? nsearch : 0; raw_out = 0; return true; }
    if (total > cap) return false;
    const uint32_t n_text = n_rows - 1;                              // the text without its '$'
    for (uint32_t r = 0; r < width; r++) {
        const uint32_t t = __ldg(ix.sa + lo + r);
        const bool inside = t >= (uint32_t)a && (uint64_t)(t - (uint32_t)a) + (uint32_t)len <= n_text;
        my_hits[incl - width + r] = inside ? t - (uint32_t)a + 1u : 0xFFFFFFFFu;
    }
    __syncwarp();
    uint32_t raw = warp_sort_unique(my_hits, total, lane);
    __syncwarp();
    if (raw && my_hits[raw - 1] == 0xFFFFFFFFu) raw--;               // the marker of the candidates that leave the text
    // ---- verification, one candidate per lane; accepted starts compacted in place (reads run ahead of the writes) ----
    uint32_t kept = 0;
    bool dirty = false;
    for (uint32_t base = 0; base < raw; base += 32) {
        const uint32_t i = base + lane;
        uint32_t v = 0;
        bool ok = false;
        if (i < raw) {
            v = my_hits[i];
            const int64_t pos = (int64_t)(v - 1u) + 200;             // padded .ref coordinate (worker_bwt.erl:53-55)
            if (!ref_plain(ix.refx, pos, len)) dirty = true;
            else ok = fs_verify_window(ix.ref + pos, Q, len, D);
        }
        __syncwarp();
        const unsigned km = __ballot_sync(FULL, ok);
        if (ok) my_hits[kept + __popc(km & ((1u << lane) - 1u))] = v;
        kept += __popc(km);
        __syncwarp();
    }
    if (__any_sync(FULL, dirty)) return false;
    c_nodes += (lane == 0) ? raw + (uint32_t)nsearch : 0;      // piece searches + verified candidates
    raw_out = kept;
    return true;
}

__global__ void __launch_bounds__(FS_WARPS * 32) fs_search_kernel(DevIndex ix, const uint8_t *__restrict__ queries,
                                                                  const int64_t *__restrict__ offs, uint32_t n_queries,
                                                                  int len, int D, uint32_t *__restrict__ hits, uint32_t cap,
                                                                  uint32_t *__restrict__ n_seeds, int32_t *__restrict__ status,
                                                                  uint4 *__restrict__ spill, int stack_cap, FsCounters *ctr, int pigeon) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FsShared &sh = *reinterpret_cast<FsShared *>(smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    uint4 *stack = sh.stack[w];
    int8_t *qc = sh.q[w];
    uint4 *my_spill = spill + ((size_t)blockIdx.x * FS_WARPS + w) * FS_SPILL;
    unsigned long long c_nodes = 0, c_lf = 0, c_raw = 0, c_seeds = 0, c_spills = 0, c_settled = 0;
    const uint32_t n_rows = ix.n;

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(&ctr->next_query, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= n_queries) break;
        const uint8_t *Q = queries + offs[qi];
        const int qlen = (int)(offs[qi + 1] - offs[qi]);
        uint32_t *my_hits = hits + (size_t)qi * cap;
        if (qlen < len + D) {                               // no usable clause / head mismatch in fs.erl
            if (lane == 0) { status[qi] = BWT_FS_SHORT; n_seeds[qi] = 0; }
            continue;
        }
        for (int k = lane; k < len + D; k += 32) qc[k] = (int8_t)bwt_code(Q[k]);
        __syncwarp();
        int top = 0, sp_top = 0;
        uint32_t raw = 0;
        bool overflow = false;
        const bool settled = pigeon && ix.ref && ix.refx && D >= 1 && D <= 2 && len > D && fs_pigeonhole(ix, Q, qc, len, D, my_hits, cap, lane, raw, c_lf, c_nodes);
        c_settled += (lane == 0 && settled) ? 1 : 0;
        // start states: the pattern ends after query character q_end = Len + (deletions - insertions)
        if (!settled) {
            raw = 0;
            const int qe = len - D + lane;
            const bool valid = lane <= 2 * D && qe >= 0;
            const unsigned vm = __ballot_sync(FULL, valid);
            if (valid) stack[__popc(vm & ((1u << lane) - 1u))] = make_uint4(0u, n_rows, pack_pq(len, qe), pack_dm(D, 0, OP_NONE));
            top = __popc(vm);
        }
        __syncwarp();

        while (top > 0 || sp_top > 0) {
            if (top == 0) {                                 // refill from the spill area
                const int take = min(sp_top, stack_cap / 2);
                for (int k = lane; k < take; k += 32) stack[k] = my_spill[sp_top - take + k];
                sp_top -= take; top = take;
                __syncwarp();
            }
            // ---- pop: lane l looks at the l-th state from the top; a prefix of them fills the warp ----
            uint4 st = make_uint4(0, 0, 0, 0);
            int cnt = 0;
            if (lane < top) {
                st = stack[top - 1 - lane];
                cnt = (st.w & 0xFFu) ? 10 : 1;
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const bool chosen = cnt > 0 && incl <= 32;
            const unsigned chosen_mask = __ballot_sync(FULL, chosen);
            const int k_states = __popc(chosen_mask);       // >= 1: the top state needs at most 10 lanes
            const int excl = incl - cnt;
            const unsigned starts = __reduce_or_sync(FULL, chosen ? (1u << excl) : 0u);
            const int total = __shfl_sync(FULL, incl, k_states - 1);
            top -= k_states;
            c_nodes += (lane == 0) ? k_states : 0;
            // ---- lane j works on child `slot` of state `owner`; the top state's children take the highest lanes ----
            const int jr = 31 - lane;
            const bool active = jr < total;
            int owner = 0, slot = 0;
            if (active) {
                const unsigned below = starts & (0xFFFFFFFFu >> (31 - jr));
                owner = __popc(below) - 1;
                const int start = 31 - __clz(below);
                slot = jr - start;
            }
            const uint32_t s_lo = __shfl_sync(FULL, st.x, owner), s_hi = __shfl_sync(FULL, st.y, owner);
            const uint32_t s_pq = __shfl_sync(FULL, st.z, owner), s_dm = __shfl_sync(FULL, st.w, owner);
            const int s_cnt = __shfl_sync(FULL, cnt, owner);
            bool push = false, accept = false;
            uint32_t nlo = 0, nhi = 0;
            uint32_t n_pq = 0, n_dm = 0;
            if (active) {
                slot = s_cnt - 1 - slot;                    // slots ascend with the lane: match lowest
                const int p = (int)(s_pq & 0xFFFFu), q = (int)(s_pq >> 16);
                const int d = (int)(s_dm & 0xFFu), m = (int)((s_dm >> 8) & 0xFFu), last = (int)((s_dm >> 16) & 0xFFu);
                const int qcode = q >= 1 ? (int)qc[q - 1] : -1;
                int np = p, nq = q, nd = d, nm = m, nlast = OP_NONE, c = -1;
                bool ok = false;
                if (slot == 0) {                            // match
                    ok = q >= 1 && qcode >= 0;
                    c = qcode; np = p - 1; nq = q - 1; nm = m < D ? m + 1 : m;
                } else if (slot <= 4) {                     // mutation to another character
                    c = slot - 1;
                    ok = q >= 1 && c != qcode;
                    np = p - 1; nq = q - 1; nd = d - 1;
                } else if (slot <= 8) {                     // insertion
                    c = slot - 5;
                    ok = !(m >= D && last == OP_DEL);
                    np = p - 1; nd = d - 1; nlast = OP_INS;
                } else {                                    // deletion
                    ok = q >= 1 && p < len && !(m >= D && last == OP_INS);
                    nq = q - 1; nd = d - 1; nlast = OP_DEL;
                }
                // can (np, nq, nd) still reach (0, 0, 0)?  nq - np = deletions - insertions to come
                const int diff = nq - np;
                ok = ok && diff <= nd && -diff <= nd;
                if (ok) {
                    if (c >= 0) { lf_extend(ix, c, s_lo, s_hi, nlo, nhi); c_lf++; }
                    else { nlo = s_lo; nhi = s_hi; }
                    if (nlo < nhi) {
                        if (np == 0) accept = nq <= nd && nd - nq <= nm;
                        else push = true;
                        n_pq = pack_pq(np, nq); n_dm = pack_dm(nd, nm, nlast);
                    }
                }
            }
            // ---- accepted intervals: locate every row ----
            unsigned acc_mask = __ballot_sync(FULL, accept);
            while (acc_mask) {
                const int src = __ffs(acc_mask) - 1;
                acc_mask &= acc_mask - 1;
                const uint32_t a_lo = __shfl_sync(FULL, nlo, src), a_hi = __shfl_sync(FULL, nhi, src);
                const uint32_t rows = a_hi - a_lo;
                if (overflow) continue;
                if (rows > cap - raw) {                     // full: the same window reached on many paths counts once
                    raw = warp_sort_unique(my_hits, raw, lane);
                    if (raw > cap / 2 || rows > cap - raw) { overflow = true; continue; }
                }
                for (uint32_t k = lane; k < rows; k += 32) my_hits[raw + k] = __ldg(ix.sa + a_lo + k) + 1u;   // N is 1-based
                raw += rows;
            }
            if (overflow) break;
            // ---- push the surviving children ----
            const unsigned push_mask = __ballot_sync(FULL, push);
            const int need = __popc(push_mask);
            if (top + need > stack_cap) {                   // spill the bottom of the stack (at least half of it)
                const int sp = max(top >> 1, top + need - stack_cap);
                if (sp_top + sp > FS_SPILL) { overflow = true; break; }
                for (int k = lane; k < sp; k += 32) my_spill[sp_top + k] = stack[k];
                sp_top += sp;
                __syncwarp();
                for (int base = sp; base < top; base += 32) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (base + lane < top) v = stack[base + lane];
                    __syncwarp();
                    if (base + lane < top) stack[base + lane - sp] = v;
                    __syncwarp();
                }
                top -= sp;
                c_spills += (lane == 0);
            }
            if (push) stack[top + __popc(push_mask & ((1u << lane) - 1u))] = make_uint4(nlo, nhi, n_pq, n_dm);
            top += need;
            __syncwarp();
        }

        if (overflow) {
            if (lane == 0) { status[qi] = BWT_FS_LIMIT; n_seeds[qi] = 0; }
            continue;
        }
        // ---- fs:find_seeds/1: sort the window starts, drop duplicates, greedy non-overlapping walk ----
        uint32_t kept = 0;
        if (raw > 0) {
            raw = warp_sort_unique(my_hits, raw, lane);
            c_raw += (lane == 0) ? raw : 0;
            if (lane == 0) {
                uint32_t cur = 1;                            // N of the window the walk stands on
                for (uint32_t k = 0; k < raw; k++) {
                    uint32_t v = my_hits[k];
                    if (v >= cur) { my_hits[kept++] = v; cur = v + (uint32_t)len; }
                }
            }
            kept = __shfl_sync(FULL, kept, 0);
        }
        if (lane == 0) { status[qi] = BWT_FS_OK; n_seeds[qi] = kept; }
        c_seeds += (lane == 0) ? kept : 0;
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) c_lf += __shfl_down_sync(FULL, c_lf, o);
    if (lane == 0) {
        atomicAdd(&ctr->nodes, c_nodes);
        atomicAdd(&ctr->lf_steps, c_lf);
        atomicAdd(&ctr->raw_hits, c_raw);
        atomicAdd(&ctr->seeds, c_seeds);
        atomicAdd(&ctr->spills, c_spills);
        atomicAdd(&ctr->settled, c_settled);
    }
}

// seeds of query r: hits[r*cap .. +n_seeds[r]) ascending  ->  out[off[r] .. off[r+1]) descending (the accumulator of
// fs:find_seeds/3), 64-bit
__global__ void fs_emit_kernel(const uint32_t *__restrict__ hits, uint32_t cap, const uint32_t *__restrict__ n_seeds,
                               const uint32_t *__restrict__ off, uint32_t n, int64_t *__restrict__ off64, int64_t base,
                               int64_t *__restrict__ out) {
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r > n) return;
    if (lane == 0) off64[r] = base + off[r];
    if (r == n) return;
    const uint32_t k = n_seeds[r];
    for (uint32_t t = lane; t < k; t += 32) out[off[r] + t] = hits[(size_t)r * cap + (k - 1 - t)];
}

}  // namespace

extern "C" int bwt_fs_find_seeds_batch(bwt_index *h, const uint8_t *queries, const int64_t *offsets, int64_t n, int32_t len,
                                       int32_t defects, int32_t max_raw_hits, int32_t *status, int64_t *seed_offsets,
                                       int64_t *seeds, int64_t cap, int64_t *total, bwt_fs_stats *stats) {
    if (!h || !offsets || n < 0 || (n > 0 && !queries) || !seed_offsets || !total) { bwt_set_error("bwt_fs_find_seeds_batch: null argument"); return BWT_ERR_ARG; }
    if (len < 2 || len > 1024 || defects < 0 || defects > FS_MAX_D) {
        bwt_set_error("bwt_fs_find_seeds_batch: Len must be 2..1024 (seeds.erl:35 needs a second argument) and Defects 0..%d", FS_MAX_D);
        return BWT_ERR_LIMIT;
    }
    if (max_raw_hits <= 0) max_raw_hits = 1024;
    int stack_cap = FS_STACK;                               // BWT_FS_STACK=32..384 shrinks the shared-memory stack (tests of the spill path)
    if (const char *e = getenv("BWT_FS_STACK")) { int v = atoi(e); if (v >= 32 && v <= FS_STACK) stack_cap = v; }
    int pigeon = 1;                                         // BWT_FS_PIGEON=0: depth-first search for every query (A/B, tests of the search)
    if (const char *e = getenv("BWT_FS_PIGEON")) pigeon = atoi(e) != 0;
    uint32_t hcap = 32;
    while (hcap < (uint32_t)max_raw_hits) hcap <<= 1;        // the sort pads to a power of two
    BWT_TRY(index_select_device(h->device));
    if (stats) memset(stats, 0, sizeof *stats);
    *total = 0;
    seed_offsets[0] = 0;
    if (n == 0) return BWT_OK;

    int dev_sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device));
    const size_t smem = sizeof(FsShared);
    CU_TRY(cudaFuncSetAttribute(fs_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fs_search_kernel, FS_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;

    // queries per pass: bounded by the scratch for raw window starts (2 GiB)
    const int64_t chunk_max = std::max<int64_t>(1, std::min<int64_t>(1 << 20, (2ll << 30) / ((int64_t)hcap * 4)));
    uint8_t *d_q = nullptr;
    int64_t *d_offs = nullptr, *d_off64 = nullptr, *d_out = nullptr;
    uint32_t *d_hits = nullptr, *d_nseeds = nullptr, *d_scan = nullptr, *d_tmp = nullptr;
    int32_t *d_status = nullptr;
    uint4 *d_spill = nullptr;
    FsCounters *d_ctr = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t stream = 0;
    int st = BWT_OK;
    int64_t base = 0;
    bool space = true;
    float ms_total = 0;
    uint32_t launches = 0;
    FsCounters sum{};
    const int64_t m_max = std::min<int64_t>(n, chunk_max);
    const unsigned grid_max = (unsigned)std::min<int64_t>((int64_t)dev_sms * per_sm, (m_max + FS_WARPS - 1) / FS_WARPS);
    do {
        if (cudaMalloc(&d_offs, (size_t)(m_max + 1) * 8) || cudaMalloc(&d_off64, (size_t)(m_max + 1) * 8) ||
            cudaMalloc(&d_hits, (size_t)m_max * hcap * 4) || cudaMalloc(&d_nseeds, (size_t)(m_max + 1) * 4) ||
            cudaMalloc(&d_scan, (size_t)(m_max + 1) * 4) || cudaMalloc(&d_tmp, scan_tmp_words((size_t)m_max + 1) * 4) ||
            cudaMalloc(&d_status, (size_t)m_max * 4) || cudaMalloc(&d_ctr, sizeof(FsCounters)) ||
            cudaMalloc(&d_spill, (size_t)grid_max * FS_WARPS * FS_SPILL * sizeof(uint4)) ||
            cudaEventCreate(&e0) || cudaEventCreate(&e1)) {
            bwt_set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_NOMEM; break;
        }
        size_t q_cap = 0, out_cap = 0;
        std::vector<int64_t> rel;
        for (int64_t lo = 0; lo < n && st == BWT_OK;) {
            const int64_t hi = std::min(n, lo + chunk_max), m = hi - lo;
            const int64_t qb = offsets[lo], bytes = offsets[hi] - qb;
            if (bytes < 0) { bwt_set_error("bwt_fs_find_seeds_batch: offsets must not decrease"); st = BWT_ERR_ARG; break; }
            if ((size_t)bytes + 16 > q_cap) {
                cudaFree(d_q); d_q = nullptr;
                q_cap = (size_t)bytes + 16;
                if (cudaMalloc(&d_q, q_cap)) { bwt_set_error("device allocation failed"); st = BWT_ERR_NOMEM; break; }
            }
            rel.resize((size_t)m + 1);
            for (int64_t k = 0; k <= m; k++) {
                rel[k] = offsets[lo + k] - qb;
                if (k && rel[k] < rel[k - 1]) { bwt_set_error("bwt_fs_find_seeds_batch: offsets must not decrease"); st = BWT_ERR_ARG; }
            }
            if (st != BWT_OK) break;
            if (bytes) cudaMemcpyAsync(d_q, queries + qb, (size_t)bytes, cudaMemcpyHostToDevice, stream);
            cudaMemcpyAsync(d_offs, rel.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, stream);
            cudaMemsetAsync(d_ctr, 0, sizeof(FsCounters), stream);
            const unsigned grid = (unsigned)std::min<int64_t>((int64_t)dev_sms * per_sm, (m + FS_WARPS - 1) / FS_WARPS);
            cudaEventRecord(e0, stream);
            fs_search_kernel<<<grid, FS_WARPS * 32, smem, stream>>>(h->dv, d_q, d_offs, (uint32_t)m, len, defects, d_hits, hcap,
                                                                    d_nseeds, d_status, d_spill, stack_cap, d_ctr, pigeon);
            cudaEventRecord(e1, stream);
            launches++;
            { cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
              if (e != cudaSuccess) { bwt_set_error("fs_search_kernel: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; } }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            ms_total += ms;
            FsCounters c;
            cudaMemcpy(&c, d_ctr, sizeof c, cudaMemcpyDeviceToHost);
            sum.nodes += c.nodes; sum.lf_steps += c.lf_steps; sum.raw_hits += c.raw_hits; sum.seeds += c.seeds; sum.spills += c.spills; sum.settled += c.settled;
            // offsets = exclusive scan of the per-query seed counts
            cudaMemcpyAsync(d_scan, d_nseeds, (size_t)m * 4, cudaMemcpyDeviceToDevice, stream);
            cudaMemsetAsync(d_scan + m, 0, 4, stream);
            if ((st = exclusive_scan_u32(d_scan, (size_t)m + 1, d_tmp, nullptr, stream, &launches)) != BWT_OK) break;
            uint32_t tot = 0;
            cudaMemcpyAsync(&tot, d_scan + m, 4, cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            if ((size_t)tot + 1 > out_cap) {
                cudaFree(d_out); d_out = nullptr;
                out_cap = (size_t)tot + 1;
                if (cudaMalloc(&d_out, out_cap * 8)) { bwt_set_error("device allocation failed"); st = BWT_ERR_NOMEM; break; }
            }
            fs_emit_kernel<<<(unsigned)((m + 1 + 7) / 8), 256, 0, stream>>>(d_hits, hcap, d_nseeds, d_scan, (uint32_t)m, d_off64, base, d_out);
            launches++;
            { cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
              if (e != cudaSuccess) { bwt_set_error("fs_emit_kernel: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; } }
            cudaMemcpy(seed_offsets + lo, d_off64, (size_t)(m + 1) * 8, cudaMemcpyDeviceToHost);
            if (status) cudaMemcpy(status + lo, d_status, (size_t)m * 4, cudaMemcpyDeviceToHost);
            if (space && seeds && base + (int64_t)tot <= cap) {
                if (tot) cudaMemcpy(seeds + base, d_out, (size_t)tot * 8, cudaMemcpyDeviceToHost);
            } else if (tot) space = false;
            base += tot;
            lo = hi;
        }
    } while (0);
    cudaFree(d_q); cudaFree(d_offs); cudaFree(d_off64); cudaFree(d_out); cudaFree(d_hits); cudaFree(d_nseeds); cudaFree(d_scan);
    cudaFree(d_tmp); cudaFree(d_status); cudaFree(d_ctr); cudaFree(d_spill);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st != BWT_OK) return st;
    *total = base;
    if (stats) {
        stats->queries = (uint64_t)n; stats->nodes = sum.nodes; stats->lf_steps = sum.lf_steps; stats->raw_hits = sum.raw_hits;
        stats->seeds = sum.seeds; stats->spills = sum.spills; stats->settled = sum.settled; stats->ms_kernel = ms_total; stats->launches = launches;
    }
    if (!space) { bwt_set_error("output capacity too small: need %lld seeds", (long long)base); return BWT_ERR_NOSPACE; }
    return BWT_OK;
}
