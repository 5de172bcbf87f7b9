// Host build of the band core (bwt_b200/csrc/swb_core.cuh) for tests/test_sw_band_rule.py: the same row update, end
// test and walk the CUDA kernel runs (swb.cu), driven from ASCII strings.  Test infrastructure only.
//   g++ -O2 -shared -fPIC -I bwt_b200/csrc tests/swb_host.cpp -o tests/_build/libswb_host.so
#include <cstdint>
#include <cstring>
#include <vector>

#include "swb_core.cuh"

namespace {

template <int WB>
int run_band(const uint8_t *q, int lq, const uint8_t *r, int lr, int p_short, int *score, char *cigar, int cap) {
    const int lo = -swb::kk_of(lq, p_short);
    swb::Band<WB> s;
    swb::init<WB>(s, lo, lr);
    std::vector<uint32_t> dirs((size_t)lq + 1);
    const int last_full = swb::imin(lq, lr - lo - WB + 1);
    for (int i = 1; i <= lq; i++) {
        uint32_t neq = 0;
        for (int b = 0; b < WB; b++) {
            const int j = i + lo + b;                      // column; its window character is r[j-1]
            const bool differ = j < 1 || j > lr || q[i - 1] != r[j - 1];
            if (differ) neq |= 1u << (2 * b);
        }
        const bool masked = i + lo < 1 || i > last_full;
        dirs[(size_t)i] = masked ? swb::row<WB, true>(s, i, lo, lr, neq) : swb::row<WB, false>(s, i, lo, lr, neq);
    }
    int b_end = 0;
    if (!swb::finish<WB>(s, lq, lo, lr, p_short, score, &b_end)) return 0;
    uint16_t runs[swb::MAX_RUNS];
    const int nr = swb::trace<WB>([&](int i) { return dirs[(size_t)i]; }, lq, lo, b_end, runs);
    if (nr < 0) return 0;
    const int clen = swb::cigar_len(runs, nr);
    if (clen + 1 > cap) return -1;
    swb::cigar_write(runs, nr, cigar);
    cigar[clen] = 0;
    return 1;
}

}  // namespace

// 1 = certified (score, cigar filled), 0 = declined (run the full table), -1 = cigar buffer too small.
// Eligibility and band width exactly as sw_band_kernel picks them; force_wb > 0 pins the width (wider bands must agree);
// p_short = the penalty budget of reads up to 128 bp (the kernel's BWT_SW_BAND_P).
extern "C" int swb_host(const uint8_t *q, int lq, const uint8_t *r, int lr, int force_wb, int p_short, int *score, char *cigar, int cap) {
    for (int i = 0; i < lq; i++) if (!memchr("ACGT", q[i], 4)) return 0;
    for (int j = 0; j < lr; j++) if (!memchr("ACGT", r[j], 4)) return 0;
    if (lq < 4 || lq > 256 || lr < lq) return 0;
    const int need = swb::need_of(lq, lr, p_short);
    int wb = need <= 6 ? 6 : need <= 8 ? 8 : need <= 12 ? 12 : need <= 16 ? 16 : 0;
    if (!wb) return 0;
    if (force_wb > 0) { if (force_wb < need) return 0; wb = force_wb; }
    switch (wb) {
        case 6: return run_band<6>(q, lq, r, lr, p_short, score, cigar, cap);
        case 8: return run_band<8>(q, lq, r, lr, p_short, score, cigar, cap);
        case 12: return run_band<12>(q, lq, r, lr, p_short, score, cigar, cap);
        case 16: return run_band<16>(q, lq, r, lr, p_short, score, cigar, cap);
    }
    return 0;
}
