// sw:sw/2 (sw.erl:47-160) for the pairs the packed-integer kernel cannot take -- an 'N' (1.6, bwt.hrl:15) or an IUPAC
// code on either side, any other byte, windows wider than the fast tiers -- one WARP per pair.
//
// The table is cut into 32 column stripes, one per lane, and the lanes run skewed: at step t lane l fills its stripe of
// row t - l + 1, takes the cell left of its stripe (value, direction, the value above it for the diagonal move and the
// running row maximum) from lane l-1 by shuffle, and hands its own last cell on -- an anti-diagonal wavefront of stripes.
// Values are IEEE doubles added in the reference's order (Vul + sigma, Vleft + gap, Vup + gap: sw.erl:90-94), one
// addition each, never contracted; ties resolve up_left > up > left (Erlang term order of {V, Atom}).  The row of
// values above a stripe lives in shared memory; directions are 2 bits per cell, one or more words per lane and step,
// written [step][word][lane] (coalesced) to the warp's slot of an HBM scratch and walked back by lane 0.  The row
// threshold (sw.erl:62,69-71) is decided by the lane that owns the last column, one row at a time, and ends the pair
// for every lane.
#include "sw.cuh"

namespace {

constexpr int GW_WARPS = 4;                 // warps (= pairs in flight) per block
#ifndef GW_MIN_BLOCKS
#define GW_MIN_BLOCKS 8                     // resident CTAs per SM the register allocation aims at: 64 registers, 32 pairs in flight per SM
                                            // (the kernel is latency-bound: 0.093 -> 0.064 us per pair against 4 CTAs x 86 registers; 12 CTAs spill: 0.076)
#endif
constexpr int GW_MAX_WC = 64;               // columns per lane at most: windows up to 2048 bases

__device__ __forceinline__ double sigma_gw(uint8_t q, uint8_t r) {      // sw.erl:164-202, first clause wins
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

// words of one warp's scratch slot for a pair: directions of every step + the CIGAR runs
__host__ __device__ inline uint64_t gw_slot_words(int lq, int lr) {
    const int wc = (lr + 31) / 32, nw = (wc + 15) / 16;
    return (uint64_t)(lq + 32) * 32 * nw + (uint64_t)lq + lr + 4;
}

__global__ void __launch_bounds__(GW_WARPS * 32, GW_MIN_BLOCKS) sw_general_warp_kernel(DevReads R, const uint8_t *__restrict__ rbase,
                                                                        const SwDesc *__restrict__ desc, const uint32_t *__restrict__ list,
                                                                        const uint32_t *__restrict__ n_list_ptr, uint32_t first, uint32_t *cursor, uint32_t *__restrict__ list_big,
                                                                        uint32_t *scratch, uint64_t slot_words, int wc_cap,
                                                                        SwOut *out, SwCounters *ctr, char *cig, uint32_t cig_cap) {
    extern __shared__ double gw_smem[];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *prev = gw_smem + (size_t)wib * wc_cap * 32;                                            // prev[c*32 + lane] = V(i-1, j)
    uint8_t *rch = reinterpret_cast<uint8_t *>(gw_smem + (size_t)GW_WARPS * wc_cap * 32) + (size_t)wib * wc_cap * 32;   // rch[c*32 + lane]
    uint32_t *slot = scratch + ((size_t)blockIdx.x * GW_WARPS + wib) * slot_words;
    const uint32_t n_list = *n_list_ptr;
    uint32_t n_done = 0;                                                    // this warp's work counters, flushed at the end
    unsigned long long n_cells = 0;
    for (;;) {
        uint32_t e = 0;
        if (lane == 0) e = first + atomicAdd(cursor, 1u);
        e = __shfl_sync(FULL, e, 0);
        if (e >= n_list) break;
        const uint32_t cand = list[e];
        const SwDesc d = desc[cand];
        const int lq = d.lq, lr = d.lr;
        const int wc = (lr + 31) / 32, nw = (wc + 15) / 16;
        if (wc > wc_cap || gw_slot_words(lq, lr) > slot_words) {            // does not fit this launch's slots: the host sizes a scratch for it
            if (lane == 0) list_big[atomicAdd(&ctr->n_big, 1u)] = cand;
            continue;
        }
        // the read's text: ASCII when it is resident (flagged reads always are), else out of the packed words
        const uint8_t *q = nullptr;
        const uint32_t *pkr = nullptr;
        if (R.ascii_full || read_is_exc(R, d.read)) { int l2; q = read_ascii(R, d.read, l2); }
        else pkr = R.pk + R.woff[d.read];
        const uint8_t *w = rbase + d.r_off;
        const int j0 = lane * wc;                                           // this lane owns columns j0+1 .. j0+mine
        const int mine = max(0, min(wc, lr - j0));
        const int last_lane = (lr - 1) / wc;                                // owner of column lr
        for (int c = 0; c < mine; c++) { prev[c * 32 + lane] = 0.0; rch[c * 32 + lane] = w[j0 + c]; }     // row 0: {0, undef}
        uint32_t pd[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};                              // directions of the row above: undef
        // what this lane hands to its right neighbour: its last cell of the row it just finished
        double o_v = 0.0, o_old = 0.0, o_max = 0.0;
        int o_dir = 3;
        const double thr0 = (double)(-2 * (lq / 2));
        int fail_row = -1;                                                  // row whose maximum fell below the threshold
        double vmax_last = 0.0;
        const int steps = lq + last_lane;
        for (int t = 0; t < steps; t++) {
            const int i = t - lane + 1;                                     // this lane's row at this step
            // from the left neighbour: its cell of the same row (it finished it one step ago)
            double l_v = __shfl_up_sync(FULL, o_v, 1), l_old = __shfl_up_sync(FULL, o_old, 1), l_max = __shfl_up_sync(FULL, o_max, 1);
            int l_dir = __shfl_up_sync(FULL, o_dir, 1);
            bool failed_here = false;
            if (i >= 1 && i <= lq && mine > 0) {
                if (lane == 0) {                                            // column 0: V = -(i+1), dir = up (sw.erl:72-77); row 0: {0, undef}
                    l_v = (double)(-(i + 1)); l_dir = 1;
                    l_old = i == 1 ? 0.0 : (double)(-i);
                    l_max = l_v;
                }
                const uint8_t qc = q ? q[i - 1] : (uint8_t)"ACGT"[read_code(pkr, lq, i - 1)];
                double left = l_v, diag = l_old, rowmax = l_max;
                int leftdir = l_dir;
                double oldv = 0.0;
#pragma unroll
                for (int x = 0; x < 4; x++) {                               // direction words: static register indices
                    if (x < nw) {
                        const uint32_t pdw = pd[x];
                        uint32_t ndw = 0u;
                        const int cend = min(16, mine - 16 * x);
                        for (int cc = 0; cc < cend; cc++) {                 // sw.erl:88-98
                            const int c = 16 * x + cc;
                            const double up = prev[c * 32 + lane];
                            const int updir = (int)((pdw >> (2 * cc)) & 3u);
                            double bv = __dadd_rn(diag, sigma_gw(qc, rch[c * 32 + lane]));
                            int bd = 2;
                            double cnd = __dadd_rn(left, leftdir == 0 ? -1.0 : -2.0);
                            if (cnd > bv) { bv = cnd; bd = 0; }
                            cnd = __dadd_rn(up, updir == 1 ? -1.0 : -2.0);
                            if (cnd > bv || (cnd == bv && bd == 0)) { bv = cnd; bd = 1; }
                            prev[c * 32 + lane] = bv;
                            ndw |= (uint32_t)bd << (2 * cc);
                            if (bv > rowmax) rowmax = bv;
                            diag = up; oldv = up; left = bv; leftdir = bd;
                        }
                        pd[x] = ndw;
                        slot[((size_t)t * nw + x) * 32 + lane] = ndw;
                    }
                }
                o_v = left; o_dir = leftdir; o_old = oldv; o_max = rowmax;
                if (lane == last_lane) {                                    // the whole row is known here: threshold of sw.erl:69-71
                    // row i is checked before row i+1 starts, and once more after the last row
                    if (rowmax < thr0 + 2.0 * (double)i) failed_here = true;
                    vmax_last = rowmax;
                }
            }
            const unsigned fm = __ballot_sync(FULL, failed_here);
            if (fm) { fail_row = t - last_lane + 1; break; }
        }
        SwOut o;
        o.score = 0; o.status = BWT_HIT_NO_MATCH; o.cigar_len = 0; o.cigar_off = 0;
        const int rows_done = fail_row >= 0 ? fail_row : lq;
        n_done++; n_cells += (unsigned long long)rows_done * (unsigned)lr;
        if (fail_row >= 0) { if (lane == 0) out[cand] = o; __syncwarp(); continue; }
        // leftmost column of the last row that holds its maximum (sw.erl:110-112); column 0 holds -(lq+1)
        const double vmax = __shfl_sync(FULL, vmax_last, last_lane);
        int jbest = 0x7FFFFFFF;
        if ((double)(-(lq + 1)) == vmax) jbest = 0;
        for (int c = mine - 1; c >= 0; c--) if (prev[c * 32 + lane] == vmax) jbest = min(jbest, j0 + c + 1);
        for (int x = 16; x > 0; x >>= 1) jbest = min(jbest, __shfl_xor_sync(FULL, jbest, x));
        __syncwarp();
        if (lane == 0) {
            // traceback (sw.erl:116-160): up_left -> M, left -> D, up -> I; runs behind the directions in the slot
            uint32_t *runs = slot + (size_t)(lq + 32) * 32 * nw;
            int i = lq, j = jbest, nr = 0, cur = -1, len = 0;
            while (i > 0) {
                int op;                                                     // 0 = M, 1 = D, 2 = I
                if (j == 0) op = 2;
                else {
                    const int l = (j - 1) / wc, c = (j - 1) - l * wc, t = i - 1 + l;
                    const uint32_t code = (slot[((size_t)t * nw + (c >> 4)) * 32 + l] >> (2 * (c & 15))) & 3u;
                    op = code == 2u ? 0 : code == 1u ? 2 : 1;
                }
                if (op == cur) len++;
                else { if (cur >= 0) runs[nr++] = ((uint32_t)len << 2) | (uint32_t)cur; cur = op; len = 1; }
                if (op == 0) { i--; j--; } else if (op == 2) i--; else j--;
            }
            runs[nr++] = ((uint32_t)len << 2) | (uint32_t)cur;
            int clen = 0;
            for (int k = 0; k < nr; k++) {
                uint32_t l = runs[k] >> 2;
                int ndg = 1;
                while (l >= 10) { ndg++; l /= 10; }
                clen += ndg + 1;
            }
            bool room;
            const uint32_t offc = sw_cig_alloc(ctr, cig_cap, (uint32_t)clen, room);
            if (room) {
                char *cp = cig + offc;
                for (int k = nr - 1; k >= 0; k--) {
                    uint32_t l = runs[k] >> 2;
                    const int op = runs[k] & 3;
                    int ndg = 1;
                    for (uint32_t x = l; x >= 10; x /= 10) ndg++;
                    for (int tt = ndg - 1; tt >= 0; tt--) { cp[tt] = (char)('0' + l % 10); l /= 10; }
                    cp[ndg] = op == 0 ? 'M' : op == 1 ? 'D' : 'I';
                    cp += ndg + 1;
                }
            }
            o.score = (int32_t)(long long)vmax;                             // trunc/1 (sw.erl:123)
            o.status = BWT_HIT_MATCH; o.cigar_len = clen; o.cigar_off = offc;
            out[cand] = o;
        }
        __syncwarp();
    }
    if (lane == 0 && n_done) {
        atomicAdd(&ctr->sw_calls, (unsigned long long)n_done);
        atomicAdd(&ctr->cells, n_cells);
        atomicAdd(&ctr->general_done, n_done);
    }
}

}  // namespace

// geometry of the always-launched general kernel for a batch whose longest read is max_lq: slots for windows up to
// max_lq + 200 + 56 bases (wider ones go to list_big), one slot per resident warp
int sw_general_warp_plan(int max_lq, int sms, SwGwPlan *pl) {
    const int lq = std::max(8, std::min(max_lq, BWT_PACK_MAX + 16));
    const int lr_cap = std::min(lq + 256, GW_MAX_WC * 32);
    pl->wc_cap = (lr_cap + 31) / 32;
    pl->slot_words = (gw_slot_words(lq, pl->wc_cap * 32) + 31) & ~31ull;
    pl->smem = (size_t)GW_WARPS * pl->wc_cap * 32 * (sizeof(double) + 1);
    pl->blocks = sms * GW_MIN_BLOCKS;
    pl->scratch_words = (uint64_t)pl->blocks * GW_WARPS * pl->slot_words;
    return BWT_OK;
}

int launch_sw_general_warp(const SwGwPlan &pl, const DevReads &R, const uint8_t *d_rbase, const SwDesc *d_desc, const uint32_t *d_list,
                           const uint32_t *d_n_list, uint32_t first, uint32_t *d_cursor, uint32_t *d_list_big, uint32_t *d_scratch, SwOut *d_out, SwCounters *d_ctr,
                           char *d_cig, uint32_t cig_cap, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set || pl.smem > 48 * 1024) {
        CU_TRY(cudaFuncSetAttribute(sw_general_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(pl.smem, 48 * 1024)));
        attr_set = true;
    }
    sw_general_warp_kernel<<<pl.blocks, GW_WARPS * 32, pl.smem, stream>>>(R, d_rbase, d_desc, d_list, d_n_list, first, d_cursor, d_list_big, d_scratch,
                                                                         pl.slot_words, pl.wc_cap, d_out, d_ctr, d_cig, cig_cap);
    K_CHECK(stream, "sw_general_warp_kernel");
    return BWT_OK;
}
