// Banded table with a certificate: sw:sw/2 (sw.erl:47-160) for pairs whose best alignment has a gap, without the
// full table.  Device code (swb.cu) and the host harness of tests/test_sw_band_rule.py compile THIS file: the test holds
// the very code the kernel runs against the tests' CPU restatement of sw:sw/2.
//
// Why a band can be exact although the reference's recurrence is not plain affine (one (V, dir) state per cell, a gap
// extended at -1 only when the neighbour's CHOSEN direction already is that gap, ties up_left > up > left):
//
//  (1) Every cell value, in the full table or in any sub-table filled by the same rule, is the plain-affine score
//      (open -2, extend -1) of the cell's own traceback path: V(c) <= A(c), the plain-affine optimum of the cell.
//  (2) Let opt = max over the last row of A, j* its leftmost column, S = the set of ALL plain-affine paths from row 0
//      (free) or column 0 to (Lq, j*) that score opt.  If no path of S has a gap longer than one character:
//        * along a path of S the reference's value equals the path's prefix score: a diagonal step gives
//          V(c) >= V(c') + sigma, a gap-open step V(c) >= V(c') - 2, and more than the prefix score would lift the end
//          cell past opt;
//        * a candidate that attains V(c) on a cell of S is the affine score of a real path to c, and no continuation of S
//          extends a gap, so that path plus the rest of S scores opt: it belongs to S.  Candidates outside S lose
//          strictly.  The chosen direction of every cell on the traceback is therefore the best-ranked step among the
//          paths of S -- in the full table and in any sub-table that holds S alike;
//        * the last row's maximum is opt, no column left of j* reaches it (V <= A < opt there): the start cell
//          (sw.erl:110-112) is (Lq, j*) in both;
//        * the row threshold -2*(Lq div 2) + 2i (sw.erl:62,69-71) holds on S while 2*Lq - opt <= 2*(Lq div 2).
//  (3) A path of penalty p = 2*Lq - score <= P stays on diagonals -k .. (Lr-Lq) + k, k = (P-1)/3: it starts on row 0
//      (diagonals 0..Lr) or in column 0, must end in the last row, and every skipped query character costs at least 3
//      (4 for the first of a run).  So the band lo = -k .. lo + WB - 1 >= (Lr-Lq) + k holds S whenever opt >= 2*Lq - P,
//      and a banded Gotoh table is exact for every (cell, state) on S.
//
// The band therefore carries two tables side by side: the reference's rule (keys 4*V + code, code left 0 / up 1 /
// up_left 2, so one integer max is lists:max/1 on {V, Dir}) and plain affine gaps in three states with a TAINT bit
// (keys 2*score + taint; a gap extension sets the bit, the max of tied keys ORs it): the bit of the end cell says whether
// ANY co-optimal path extends a gap.  Clean and within budget: the band's traceback is the reference's result.
// Anything else runs the full table.  Evidence beside the argument: tests/test_sw_band_rule.py (CPU, millions of
// adversarial pairs during development) and tests/test_gpu_parity.py (the kernel, every SW / align test).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SWB_HD __host__ __device__ __forceinline__
#define SWB_UNROLL _Pragma("unroll")
#else
#define SWB_HD inline
#define SWB_UNROLL
#endif

namespace swb {

constexpr int NEGK = -(1 << 20);
constexpr int MAX_RUNS = 40;                 // CIGAR runs a certified path can have: at most 2 per gap + 1, gaps <= P/2

SWB_HD int imax(int a, int b) { return a > b ? a : b; }
SWB_HD int imin(int a, int b) { return a < b ? a : b; }

// penalty budget by read length (the band's half width grows with it) and the row-threshold limit 2*(Lq div 2);
// p_short = the budget of reads up to 128 bp (9: one indel and a substitution, band of Dd + 5 diagonals; 12: one indel and
// two substitutions, Dd + 7) -- any budget is sound, it trades cells per pair against pairs left to the full table
SWB_HD int pmax_of(int lq, int p_short) { return imin(lq <= 128 ? p_short : lq <= 192 ? imax(12, p_short) : 15, 2 * (lq >> 1)); }
SWB_HD int kk_of(int lq, int p_short) { return imax(0, (pmax_of(lq, p_short) - 1) / 3); }
// diagonals the band must hold: -k .. dd + k
SWB_HD int need_of(int lq, int lr, int p_short) { return (lr - lq) + 2 * kk_of(lq, p_short) + 1; }

template <int WB>
struct Band {
    int K[WB + 1];       // reference rule: 4*V + code
    int H[WB + 1];       // plain affine, best of the three states: 2*score + taint
    int I[WB + 1];       // plain affine, state "query character skipped" (up)
};

// row 0: {0, undef} on columns 0..Lr (sw.erl:61); band index b <-> column i + lo + b
template <int WB>
SWB_HD void init(Band<WB> &s, int lo, int lr) {
SWB_UNROLL
    for (int b = 0; b <= WB; b++) {
        const int j = lo + b;
        const bool in = b < WB && j >= 0 && j <= lr;
        s.K[b] = in ? 2 : NEGK + 2;          // code up_left: no gap bonus for the cells below / right of row 0
        s.H[b] = in ? 0 : NEGK;
        s.I[b] = NEGK;
    }
}

// row i (1-based) in place; neq bit 2b set <=> query character i differs from window character i + lo + b (1-based).
// Returns the row's direction codes, 2 bits per band index.  MASKED rows may touch columns <= 0 or > Lr.
template <int WB, bool MASKED>
SWB_HD uint32_t row(Band<WB> &s, int i, int lo, int lr, uint32_t neq) {
    int kl = NEGK + 2, hl = NEGK, dl = NEGK;                 // left of band index 0: outside the band
    uint32_t drow = 0;
SWB_UNROLL
    for (int b = 0; b < WB; b++) {
        const bool mm = (neq >> (2 * b)) & 1u;
        const int ku_src = s.K[b + 1];
        const int kd = (s.K[b] & ~3) + (mm ? -4 : 8) + 2;                     // sw.erl:91
        const int ku = (ku_src & ~3) + ((ku_src & 3) == 1 ? -3 : -7);         // get_Uvalue: -1 behind an `up`, else -2
        const int kf = (kl & ~3) + ((kl & 3) == 0 ? -4 : -8);                 // get_Lvalue: -1 behind a `left`, else -2
        int kc = imax(kd, imax(ku, kf));
        const int m = s.H[b] + (mm ? -2 : 4);
        int ic = imax(s.H[b + 1] - 4, (s.I[b + 1] - 2) | 1);
        int dc = imax(hl - 4, (dl - 2) | 1);
        int hc = imax(m, imax(ic, dc));
        if (MASKED) {
            const int j = i + lo + b;
            if (j < 0 || j > lr) { kc = NEGK + 2; hc = NEGK; ic = NEGK; dc = NEGK; }
            else if (j == 0) {                                                // column 0: {-(i+1), up} (sw.erl:72-77)
                kc = 4 * (-(i + 1)) + 1;
                hc = ic = 2 * (-(i + 1)) + (i >= 2 ? 1 : 0);                  // a query-side gap of i characters
                dc = NEGK;
            }
        }
        s.K[b] = kc; s.H[b] = hc; s.I[b] = ic;
        kl = kc; hl = hc; dl = dc;
        drow |= (uint32_t)(kc & 3) << (2 * b);
    }
    return drow;
}

// after row Lq: the leftmost best end of the affine table; certified <=> within the budget, untainted, and the
// reference-rule table agrees on the value (it must: kept as a guard)
template <int WB>
SWB_HD bool finish(const Band<WB> &s, int lq, int lo, int lr, int p_short, int *score, int *b_end) {
    int best = NEGK, bb = -1;
SWB_UNROLL
    for (int b = 0; b < WB; b++) {
        const int j = lq + lo + b;
        const int sc = s.H[b] >> 1;
        if (j >= 0 && j <= lr && sc > best) { best = sc; bb = b; }
    }
    if (bb < 0) return false;
    int hk = NEGK, kk = NEGK;
SWB_UNROLL
    for (int b = 0; b < WB; b++) if (b == bb) { hk = s.H[b]; kk = s.K[b]; }
    if (2 * lq - best > pmax_of(lq, p_short)) return false;
    if (hk & 1) return false;
    if ((kk >> 2) != best) return false;
    *score = best; *b_end = bb;
    return true;
}

// walk back from (Lq, b); dir_at(i) = the direction word of row i.  runs[k] = (len << 2) | op, op 0 M / 1 D / 2 I, LAST
// run first.  Returns the number of runs, or -1 when the walk leaves the band or the runs do not fit.
template <int WB, class DirAt>
SWB_HD int trace(DirAt dir_at, int lq, int lo, int b, uint16_t *runs) {
    int i = lq, nr = 0, cur = -1, len = 0;
    int guard = 2 * lq + 4 * WB + 8;
    while (i > 0) {
        if (b < 0 || b >= WB || --guard < 0) return -1;
        const int j = i + lo + b;
        if (j < 0) return -1;
        const uint32_t code = j == 0 ? 1u : (dir_at(i) >> (2 * b)) & 3u;
        const int op = code == 2u ? 0 : code == 1u ? 2 : 1;
        if (op == cur) len++;
        else {
            if (cur >= 0) { if (nr >= MAX_RUNS) return -1; runs[nr++] = (uint16_t)((len << 2) | cur); }
            cur = op; len = 1;
        }
        if (op == 0) i--;
        else if (op == 2) { i--; b++; }
        else b--;
    }
    if (nr >= MAX_RUNS) return -1;
    runs[nr++] = (uint16_t)((len << 2) | cur);
    return nr;
}

SWB_HD int cigar_len(const uint16_t *runs, int nr) {
    int clen = 0;
    for (int k = 0; k < nr; k++) {
        const int l = (int)(runs[k] >> 2);
        clen += 2 + (l >= 10) + (l >= 100) + (l >= 1000);
    }
    return clen;
}

SWB_HD void cigar_write(const uint16_t *runs, int nr, char *c) {
    for (int k = nr - 1; k >= 0; k--) {
        const int v = (int)runs[k];
        int l = v >> 2;
        const int op = v & 3;
        const int nd = 1 + (l >= 10) + (l >= 100) + (l >= 1000);
        for (int t = nd - 1; t >= 0; t--) { c[t] = (char)('0' + l % 10); l /= 10; }
        c[nd] = op == 0 ? 'M' : op == 1 ? 'D' : 'I';
        c += nd + 1;
    }
}

}  // namespace swb
