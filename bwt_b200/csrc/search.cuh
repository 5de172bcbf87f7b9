// Device-side FM backward search with the reference's exact (bug-for-bug) semantics.
//
//   fmi:se_down/4, fmi:se_up/3   fmi.erl:25-39    -> lf_step()   (closed form)
//   bwa:find_seeds/6,5           bwa.erl:12-45    -> search_sub()
//
// The reference stores the 0-based LF target I in every row and uses it as a 1-based row number, and
// se_down excludes Ep while se_up includes it.  With occ(c,e) = #{rows e' <= e : L[e'] = c} and P_c the
// 1-based first row whose F is c:
//      found  <=>  Sp < Ep  and  occ(c,Ep-1) > occ(c,Sp-1)
//      Sp' = P_c - 1 + occ(c,Sp-1)          Ep' = P_c - 2 + occ(c,Ep)
// occ(c,Ep-1) = occ(c,Ep) - [L[Ep] = c], so one step needs two 32-byte Occ blocks (one when Sp-1 and Ep
// fall into the same block).
#pragma once
#include "common.cuh"

struct OccHit {
    uint32_t occ;     // occ(c, e)
    uint32_t is_c;    // L[e] == c
};

// occ(c,e) for 1 <= e <= n from one 32-byte block
__device__ __forceinline__ OccHit occ_at(const DevIndex &ix, uint32_t e, int c) {
    uint32_t k = e - 1;
    uint32_t blk = k >> 6, r = k & 63;          // symbol k sits at bit r of the block's planes
    const uint4 *p = ix.occ + 2 * (size_t)blk;
    uint4 cnt = __ldg(p);
    uint4 pl = __ldg(p + 1);
    uint32_t base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
    uint32_t lo0 = (c & 1) ? pl.x : ~pl.x, lo1 = (c & 1) ? pl.y : ~pl.y;
    uint32_t hi0 = (c & 2) ? pl.z : ~pl.z, hi1 = (c & 2) ? pl.w : ~pl.w;
    uint32_t m0 = lo0 & hi0, m1 = lo1 & hi1;    // match bits of symbols 0..31 / 32..63
    // symbols 0..r inclusive
    uint32_t cnt_in;
    uint32_t bit;
    if (r < 32) {
        cnt_in = __popc(m0 & (0xFFFFFFFFu >> (31 - r)));
        bit = (m0 >> r) & 1u;
    } else {
        cnt_in = __popc(m0) + __popc(m1 & (0xFFFFFFFFu >> (63 - r)));
        bit = (m1 >> (r - 32)) & 1u;
    }
    OccHit o;
    o.occ = base + cnt_in;
    o.is_c = bit;
    if (c == 0) {                                // the '$' row is stored as A
        o.occ -= (ix.dollar <= e) ? 1u : 0u;
        if (e == ix.dollar) o.is_c = 0;
    }
    return o;
}

// One step of bwa.erl:36-43.  Returns false when se_down says not_found.
__device__ __forceinline__ bool lf_step(const DevIndex &ix, uint32_t &sp, uint32_t &ep, int c) {
    if (sp >= ep) return false;                  // se_down(Ep,Ep,_,_) -> not_found  (fmi.erl:25)
    uint32_t a = (sp > 1) ? occ_at(ix, sp - 1, c).occ : 0u;
    OccHit hi = occ_at(ix, ep, c);
    uint32_t b = hi.occ - hi.is_c;               // occ(c, Ep-1)
    if (b <= a) return false;
    // select chain, not ix.P[c]: an indexed read makes the compiler keep a 104-byte copy of the kernel parameter in local memory
    const uint32_t pc = c == 0 ? ix.P[0] : c == 1 ? ix.P[1] : c == 2 ? ix.P[2] : ix.P[3];
    sp = pc - 1 + a;
    ep = pc - 2 + hi.occ;
    return true;
}

struct SearchRes {
    uint32_t sp, ep;
    int n;        // characters matched (the reference's N)
    int status;   // BWT_SEEDS / BWT_NO_SEEDS / BWT_TOO_MANY_SEEDS / BWT_SEARCH_CRASH
    int table;    // characters matched by a k-mer table lookup (0 = none): their LF steps touched no Occ block
};

__device__ __forceinline__ void first_interval(const DevIndex &ix, int c, uint32_t &sp, uint32_t &ep) {
    // bwa.erl:15-24 ('N' is treated like 'A' by the caller)
    if (c == 0) { sp = 1; ep = ix.P[1] - 1; }
    else if (c == 1) { sp = ix.P[1]; ep = ix.P[2] - 1; }
    else if (c == 2) { sp = ix.P[2]; ep = ix.P[3] - 1; }
    else { sp = ix.P[3]; ep = ix.last; }
}

// entry of the depth-`depth` interval table (8 <= depth <= ix.deep_k) for the k-mer idx (last character in bits 0-1)
__device__ __forceinline__ uint2 kmer_lookup(const DevIndex &ix, int depth, uint32_t idx) {
    idx &= (uint32_t)((1ull << (2 * depth)) - 1ull);
    if (depth == BWT_MAX_LEN)
        return ix.fused ? __ldg(reinterpret_cast<const uint2 *>(ix.fused + 4 * (size_t)idx)) : __ldg(ix.kdeep + idx);
    return __ldg(ix.ktab + ktab_off(depth) + idx);
}

// bwa:find_seeds/6 on q[0..len), len >= 1, once the characters next to the end are known as 2-bit codes:
// c = code of the last character ('N' counts as 'A'), idx = codes of the last `good` characters (last one in
// bits 0-1), all of them plain ACGT.  use_table: consult the k-mer tables (false while building them).
// PACKED: q is null and the characters come from the packed read pkr (length lq_total) -- plain ACGT throughout.
template <bool PACKED>
__device__ __forceinline__ SearchRes search_core_t(const DevIndex &ix, const uint8_t *q, const uint32_t *pkr, int lq_total, int len,
                                                   int c, uint32_t idx, int good, bool use_table) {
    SearchRes R;
    R.table = 0;
    uint32_t sp = 0, ep = 0;
    int n = 1, k = len - 2;
    bool have = false, stopped = false;
    if (use_table && good >= BWT_MIN_LEN) {
        // the table of depth d applies when the d-1 characters before the last are plain ACGT: the deepest one that fits
        const int depth = good < ix.deep_k ? good : ix.deep_k;
        const uint2 e = kmer_lookup(ix, depth, idx);
        int ne = (int)(e.y >> 28);
        if (ne) {
            R.table = ne;                              // characters whose steps the table served
            sp = e.x; ep = e.x + (e.y & KMER_W_MASK); n = ne; k = len - 1 - ne;
            have = true;
            stopped = ne < depth;                      // not_found inside the table's reach
        }
    }
    if (!have) first_interval(ix, c, sp, ep);
    bool checked = false;
    while (!stopped) {
        if (n == BWT_MAX_LEN) { checked = true; break; }             // bwa.erl:31-32
        if (k < 0) { checked = true; break; }                        // bwa.erl:44-45
        int cc;
        if (PACKED) cc = read_code(pkr, lq_total, k);
        else {
            uint8_t ch = q[k];
            if (ch == 'N') { checked = true; break; }                // bwa.erl:33-34
            cc = bwt_code(ch);
            if (cc < 0) {
                if (ch == '$' && sp < ep && ix.dollar >= sp && ix.dollar <= ep - 1) {
                    // the '$' row's I is 0: the walk lands on (0,0) and every later clause reaches
                    // get_seed_ends(FM,0,0,N): no_seeds for N < 8, element(0,FM) = badarg otherwise
                    n++;
                    R.sp = R.ep = 0; R.n = n; R.status = n < BWT_MIN_LEN ? BWT_NO_SEEDS : BWT_SEARCH_CRASH;
                    return R;
                }
                checked = false; break;                              // no row has this L: not_found
            }
        }
        if (!lf_step(ix, sp, ep, cc)) { checked = false; break; }    // bwa.erl:37-38: no width check here
        n++; k--;
    }
    R.sp = sp; R.ep = ep; R.n = n;
    if (checked && ep - sp > BWT_MAX_RANGE) R.status = BWT_TOO_MANY_SEEDS;
    else R.status = n >= BWT_MIN_LEN ? BWT_SEEDS : BWT_NO_SEEDS;
    return R;
}

__device__ __forceinline__ SearchRes search_core(const DevIndex &ix, const uint8_t *q, int len, int c, uint32_t idx,
                                                 int good, bool use_table) {
    return search_core_t<false>(ix, q, nullptr, 0, len, c, idx, good, use_table);
}

// bwa:find_seeds/6 on the first `end` bases of a packed read (plain ACGT): idx = its last 15 (or fewer) codes
__device__ __forceinline__ SearchRes search_packed(const DevIndex &ix, const uint32_t *pkr, int lq, int end, uint32_t idx) {
    return search_core_t<true>(ix, nullptr, pkr, lq, end, (int)(idx & 3u), idx, end < BWT_MAX_LEN ? end : BWT_MAX_LEN, true);
}

// bwa:find_seeds/6 on q[0..len), len >= 1, from the ASCII bytes
__device__ __forceinline__ SearchRes search_sub(const DevIndex &ix, const uint8_t *q, int len, bool use_table) {
    uint8_t h = q[len - 1];
    int c = (h == 'N') ? 0 : bwt_code(h);
    if (c < 0) {                                       // case_clause (bwa.erl:15-24)
        SearchRes R;
        R.sp = R.ep = 0; R.n = 0; R.status = BWT_SEARCH_CRASH; R.table = 0;
        return R;
    }
    uint32_t idx = (uint32_t)c;
    int good = 1;                                      // characters from the end that are plain ACGT
    if (use_table) {
        const int want = len < BWT_MAX_LEN ? len : BWT_MAX_LEN;
        for (int t = 1; t < want; t++) {
            int ct = bwt_code(q[len - 1 - t]);
            if (ct < 0) break;
            idx |= (uint32_t)ct << (2 * t);
            good++;
        }
    }
    return search_core(ix, q, len, c, idx, good, use_table);
}
