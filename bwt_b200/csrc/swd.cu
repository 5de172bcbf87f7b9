// sw:sw/2 (sw.erl:47-160) in IEEE doubles, one THREAD per pair: the throughput kernel for pairs that hold an 'N' (1.6,
// bwt.hrl:15) or an IUPAC code.  The warp-per-pair kernel (swg.cu) is a latency design -- one pair costs it ~0.2 ms of
// dependent steps and it keeps 32 pairs in flight per SM -- which is right for worker-sized batches and wrong for a
// batch in which two thirds of the reads carry an 'N'.  Here every thread walks its own table the way sw_fast_kernel
// does: strips of SWD_R rows in registers (value, direction, running row maximum), the strip's boundary row (value +
// direction per column) and the 2-bit directions in a per-thread slice of an HBM / L2 scratch laid out [column][thread]
// (coalesced), strips aligned at the LAST row so that the boundary row after the last strip is the row the start cell is
// searched in.  Values are doubles added in the reference's order, one addition each, never contracted (-fmad=false);
// ties resolve up_left > up > left (Erlang term order of {V, Atom}, sw.erl:90-94).
//
// sigma/2 (sw.erl:164-202) per cell is a table look-up: the window byte's class (A C G T N R Y K M S W B D V) comes from a
// 256-entry constant table, the query character's row of 2-bit codes (-1 / 2 / 1.6 per class) is built once per strip.
// A window byte outside those fourteen letters matches only itself, which no table of classes can say: such pairs, and
// pairs beyond the slice's columns or strips, go on to the warp-per-pair kernel (list `next`).
#include <type_traits>

#include "sw.cuh"

namespace {

constexpr int SWD_R = 8;                    // rows per register strip: 8 x 2 bits of directions = one uint16 per column
constexpr int SWD_BLOCK = 128;
constexpr int SWD_OTHER = 15;
#ifndef SWD_MIN_BLOCKS
#define SWD_MIN_BLOCKS 5                    // resident CTAs per SM the register allocation aims at (3: 0.027, 4: 0.025, 5: 0.024 us per pair)
#endif

__device__ __forceinline__ double swd_sigma(uint8_t q, uint8_t r) {       // sw.erl:164-202, first clause wins
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

__device__ __forceinline__ int swd_class(uint8_t c) {
    switch (c) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4;
    case 'R': return 5; case 'Y': return 6; case 'K': return 7; case 'M': return 8; case 'S': return 9;
    case 'W': return 10; case 'B': return 11; case 'D': return 12; case 'V': return 13;
    default: return SWD_OTHER;
    }
}

// the query character's row of the substitution table: 2 bits per window class, 0 = -1, 1 = 2, 2 = 1.6
__device__ __forceinline__ uint32_t swd_row_word(uint8_t q) {
    const char rep[15] = "ACGTNRYKMSWBDV";
    uint32_t w = 0;
#pragma unroll
    for (int c = 0; c < 14; c++) {
        const double s = swd_sigma(q, (uint8_t)rep[c]);
        w |= (s == 2.0 ? 1u : s == 1.6 ? 2u : 0u) << (2 * c);
    }
    return w;
}

struct SwdParams {
    DevReads R;
    const uint8_t *rbase;
    const SwDesc *desc;
    const uint32_t *list;
    const uint32_t *n_list;   // device counter: entries of list
    uint32_t *next;           // pairs this kernel does not take ...
    uint32_t *n_next;         // ... and their count
    SwOut *out;
    SwCounters *ctr;
    char *cig;
    uint32_t cig_cap;
    double *bval;             // boundary row: bval[j * NT + t]
    uint8_t *bdir;            //               bdir[j * NT + t]
    uint16_t *dirs;           // dirs[(s * cap_cols + j) * NT + t]: rows r of strip s at bits 2r
    int cap_cols;             // columns (Lr + 1) a slice holds
    int max_strips;
};

// direction codes as swg.cu: 0 left, 1 up, 2 up_left, 3 undef (row 0)
__global__ void __launch_bounds__(SWD_BLOCK, SWD_MIN_BLOCKS) sw_double_kernel(SwdParams p) {
    constexpr int R = SWD_R;
    const size_t NT = (size_t)gridDim.x * SWD_BLOCK;
    const size_t gt = (size_t)blockIdx.x * SWD_BLOCK + threadIdx.x;
    double *bval = p.bval + gt;
    uint8_t *bdir = p.bdir + gt;
    uint16_t *dirs = p.dirs + gt;
    const uint32_t n_list = *p.n_list;
    unsigned long long calls = 0, cells = 0;
    for (size_t e = gt; e < n_list; e += NT) {
        const uint32_t cand = p.list[e];
        const SwDesc d = p.desc[cand];
        const int lq = d.lq, lr = d.lr;
        const uint8_t *w = p.rbase + d.r_off;
        const int nstr = (lq + R - 1) / R;
        bool take = lq >= 1 && lr >= 1 && lr + 1 <= p.cap_cols && nstr <= p.max_strips;
        if (take) {
            int other = 0;
            for (int j = 0; j < lr; j++) other |= swd_class(w[j]) == SWD_OTHER;
            take = !other;
        }
        if (!take) { p.next[atomicAdd(p.n_next, 1u)] = cand; continue; }
        // the read's text: ASCII when it is resident (flagged reads always are), else out of the packed words
        const uint8_t *q = nullptr;
        const uint32_t *pkr = nullptr;
        if (p.R.ascii_full || read_is_exc(p.R, d.read)) { int l2; q = read_ascii(p.R, d.read, l2); }
        else pkr = p.R.pk + p.R.woff[d.read];
        const int i_first = lq - nstr * R + 1;                               // row of strip 0, r = 0 (<= 1): rows <= 0 act as row 0
        for (int j = 1; j <= lr; j++) { bval[(size_t)j * NT] = 0.0; bdir[(size_t)j * NT] = 3; }   // row 0: {0, undef} (sw.erl:61)
        const double thr0 = (double)(-2 * (lq / 2));                         // sw.erl:62
        int fail_row = -1;
        bool handed_on = false;
        double slack = 0.0;                                                  // row maximum - threshold at the last row of the previous strip
        for (int s = 0; s < nstr && fail_row < 0; s++) {
            const int i0 = i_first + s * R;
            uint32_t rowtab[R];
            double vl[R], mx[R];
            uint32_t dl = 0;                                                 // directions of the cells to the left, 2 bits per row
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = i0 + r;
                const uint8_t qc = i >= 1 ? (q ? q[i - 1] : (uint8_t)"ACGT"[read_code(pkr, lq, i - 1)]) : (uint8_t)'A';
                rowtab[r] = swd_row_word(qc);
                // column 0: V = -(i+1), dir = up (sw.erl:72-77); rows <= 0: {0, undef}
                vl[r] = i >= 1 ? (double)(-(i + 1)) : 0.0;
                mx[r] = vl[r];
                dl |= (i >= 1 ? 1u : 3u) << (2 * r);
            }
            double top_prev = i0 - 1 >= 1 ? (double)(-i0) : 0.0;             // V(i0-1, 0)
            uint16_t *dstrip = dirs + (size_t)s * p.cap_cols * NT;
            // A row's maximum is at most 2 below the maximum of the row above (the `up` candidate under the best cell) and
            // the threshold rises by 2 per row: with a slack of more than 4 per row of this strip no row of it can fail
            // (+1: a subtraction that crosses a binade may round in the last place), and only the strip's last row needs
            // its maximum.  Rows <= 0 exist in the first strip only.
            const bool masked = i0 < 1, light = !masked && slack >= 4.0 * R + 1.0;
            auto columns = [&](auto MASKED, auto TRACK) {
                for (int j = 1; j <= lr; j++) {
                    const double top = bval[(size_t)j * NT];                 // V(i0-1, j)
                    int updir = (int)bdir[(size_t)j * NT];
                    const int sh = 2 * swd_class(w[j - 1]);
                    double diag = top_prev, up = top;
                    uint32_t nd = 0;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t code = (rowtab[r] >> sh) & 3u;
                        const double sg = code == 1u ? 2.0 : code == 2u ? 1.6 : -1.0;
                        const double dg = __dadd_rn(diag, sg);               // sw.erl:90-94: one addition per candidate
                        const double uc = __dadd_rn(up, updir == 1 ? -1.0 : -2.0);
                        const double lc = __dadd_rn(vl[r], ((dl >> (2 * r)) & 3u) == 0u ? -1.0 : -2.0);
                        const double v1 = fmax(dg, uc);                      // up_left > up > left among equal values
                        int dir = uc > dg ? 1 : 2;
                        double v = fmax(v1, lc);
                        if (lc > v1) dir = 0;
                        if (decltype(MASKED)::value && i0 + r < 1) { v = 0.0; dir = 3; }   // a row of the padding in front of the first strip
                        diag = vl[r];
                        vl[r] = v;
                        if (decltype(TRACK)::value || r == R - 1) mx[r] = fmax(mx[r], v);
                        nd |= (uint32_t)dir << (2 * r);
                        up = v; updir = dir;
                    }
                    dl = nd;
                    top_prev = top;
                    bval[(size_t)j * NT] = up;                               // the strip's last row is the next strip's boundary
                    bdir[(size_t)j * NT] = (uint8_t)updir;
                    dstrip[(size_t)j * NT] = (uint16_t)nd;
                }
            };
            if (masked) columns(std::true_type{}, std::true_type{});
            else if (light) columns(std::false_type{}, std::false_type{});
            else columns(std::false_type{}, std::true_type{});
            // row threshold (sw.erl:69-71): the maximum of row i must reach thr0 + 2*i
            if (!light) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = i0 + r;
                    if (i >= 1 && fail_row < 0 && mx[r] < thr0 + 2.0 * (double)i) fail_row = i;
                }
            }
            slack = mx[R - 1] - (thr0 + 2.0 * (double)(i0 + R - 1));
            if (s == nstr - 1 && fail_row < 0) {
                // leftmost column of the last row that holds its maximum (sw.erl:110-112); column 0 holds -(lq+1)
                const double vmax = mx[R - 1];
                int jbest = -1;
                if ((double)(-(lq + 1)) == vmax) jbest = 0;
                for (int j = 1; j <= lr && jbest < 0; j++) if (bval[(size_t)j * NT] == vmax) jbest = j;
                // traceback (sw.erl:116-160): up_left -> M, left -> D, up -> I; the runs go where the boundary row was
                uint32_t *runs = reinterpret_cast<uint32_t *>(p.bval) + 2 * gt;        // runs[k * 2 * NT]: the low word of this thread's bval[k]
                int i = lq, j = jbest, nr = 0, cur = -1, len = 0;
                const int run_cap = p.cap_cols;                              // one run per boundary-row slot
                while (i > 0) {
                    int op;                                                  // 0 = M, 1 = D, 2 = I
                    if (j == 0) op = 2;
                    else {
                        const int t = i - i_first, ss = t / R, r = t - ss * R;
                        const uint32_t c = ((uint32_t)dirs[((size_t)ss * p.cap_cols + j) * NT] >> (2 * r)) & 3u;
                        op = c == 2u ? 0 : c == 1u ? 2 : 1;
                    }
                    if (op == cur) len++;
                    else { if (cur >= 0) { if (nr < run_cap) runs[(size_t)nr * 2 * NT] = ((uint32_t)len << 2) | (uint32_t)cur; nr++; } cur = op; len = 1; }
                    if (op == 0) { i--; j--; } else if (op == 2) i--; else j--;
                }
                if (nr < run_cap) runs[(size_t)nr * 2 * NT] = ((uint32_t)len << 2) | (uint32_t)cur;
                nr++;
                if (nr > run_cap) {                                          // more runs than slots (a path that alternates all the way): the warp kernel's
                    p.next[atomicAdd(p.n_next, 1u)] = cand;
                    handed_on = true;
                    break;
                }
                int clen = 0;
                for (int k = 0; k < nr; k++) {
                    uint32_t l = runs[(size_t)k * 2 * NT] >> 2;
                    int ndg = 1;
                    while (l >= 10) { ndg++; l /= 10; }
                    clen += ndg + 1;
                }
                bool room;
                const uint32_t offc = sw_cig_alloc(p.ctr, p.cig_cap, (uint32_t)clen, room);
                if (room) {
                    char *cp = p.cig + offc;
                    for (int k = nr - 1; k >= 0; k--) {
                        uint32_t l = runs[(size_t)k * 2 * NT] >> 2;
                        const int op = (int)(runs[(size_t)k * 2 * NT] & 3u);
                        int ndg = 1;
                        for (uint32_t x = l; x >= 10; x /= 10) ndg++;
                        for (int tt = ndg - 1; tt >= 0; tt--) { cp[tt] = (char)('0' + l % 10); l /= 10; }
                        cp[ndg] = op == 0 ? 'M' : op == 1 ? 'D' : 'I';
                        cp += ndg + 1;
                    }
                }
                SwOut o;
                o.score = (int32_t)(long long)vmax;                          // trunc/1 (sw.erl:123)
                o.status = BWT_HIT_MATCH; o.cigar_len = clen; o.cigar_off = offc;
                p.out[cand] = o;
            }
        }
        if (handed_on) continue;
        if (fail_row >= 0) {
            SwOut o;
            o.score = 0; o.status = BWT_HIT_NO_MATCH; o.cigar_len = 0; o.cigar_off = 0;
            p.out[cand] = o;
        }
        calls++;
        cells += (unsigned long long)(fail_row >= 0 ? fail_row : lq) * (unsigned)lr;
    }
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
        calls += __shfl_down_sync(0xFFFFFFFFu, calls, dlt);
        cells += __shfl_down_sync(0xFFFFFFFFu, cells, dlt);
    }
    if ((threadIdx.x & 31) == 0 && calls) {
        atomicAdd(&p.ctr->sw_calls, calls);
        atomicAdd(&p.ctr->cells, cells);
        atomicAdd(&p.ctr->general_done, (uint32_t)calls);
    }
}

}  // namespace

// geometry for a batch whose longest read is max_lq: windows up to max_lq + 176 columns, grids of up to 8 CTAs per SM,
// fewer when the slices would not fit 3 GiB of scratch
int sw_double_plan(int max_lq, int sms, uint32_t n_cap, SwdPlan *pl) {
    const int lq = std::max(8, std::min(max_lq, 1040));
    pl->cap_cols = ((lq + 176) + 15) & ~15;
    pl->max_strips = (lq + SWD_R - 1) / SWD_R;
    const uint64_t per_thread = (uint64_t)pl->cap_cols * 9 + (uint64_t)pl->max_strips * pl->cap_cols * 2;
    uint64_t blocks = std::min<uint64_t>((uint64_t)sms * 8, ((uint64_t)n_cap + SWD_BLOCK - 1) / SWD_BLOCK);
    blocks = std::min<uint64_t>(blocks, std::max<uint64_t>(1, (3ull << 30) / (per_thread * SWD_BLOCK)));
    pl->blocks = (int)std::max<uint64_t>(1, blocks);
    const uint64_t nt = (uint64_t)pl->blocks * SWD_BLOCK;
    pl->off_bdir = nt * pl->cap_cols * 8;
    pl->off_dirs = (pl->off_bdir + nt * pl->cap_cols + 15) & ~15ull;
    pl->scratch_bytes = pl->off_dirs + nt * pl->max_strips * pl->cap_cols * 2;
    return BWT_OK;
}

int launch_sw_double(const SwdPlan &pl, const DevReads &R, const uint8_t *d_rbase, const SwDesc *d_desc, const uint32_t *d_list,
                     const uint32_t *d_n_list, uint32_t *d_next, uint32_t *d_n_next, uint8_t *d_scratch, SwOut *d_out, SwCounters *d_ctr,
                     char *d_cig, uint32_t cig_cap, cudaStream_t stream) {
    SwdParams p;
    p.R = R; p.rbase = d_rbase; p.desc = d_desc; p.list = d_list; p.n_list = d_n_list; p.next = d_next; p.n_next = d_n_next;
    p.out = d_out; p.ctr = d_ctr; p.cig = d_cig; p.cig_cap = cig_cap;
    p.bval = reinterpret_cast<double *>(d_scratch);
    p.bdir = d_scratch + pl.off_bdir;
    p.dirs = reinterpret_cast<uint16_t *>(d_scratch + pl.off_dirs);
    p.cap_cols = pl.cap_cols; p.max_strips = pl.max_strips;
    sw_double_kernel<<<pl.blocks, SWD_BLOCK, 0, stream>>>(p);
    K_CHECK(stream, "sw_double_kernel");
    return BWT_OK;
}
