// Seeding kernels (see seed.cuh).  Reference semantics: sga.erl:17-76, bwa.erl:12-45, fmi.erl:25-39,89-100.
#include <cooperative_groups.h>
#include <cstdlib>

#include "search.cuh"
#include "seed.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SGA_BLOCK = 128;
#ifndef SGA_MIN_BLOCKS
#define SGA_MIN_BLOCKS 10      // resident CTAs per SM the register allocation of the 16-lane kernel aims at: 51 registers, 40 warps per
                               // SM (7.72 -> 7.21 ms per 10 M reads against 8 CTAs x 64 registers; 12 would spill: 7.49 ms).  The
                               // 32-lane kernel's shared memory allows 8.
#endif

enum { CTR_SEARCHES = 0, CTR_LF, CTR_HBM_LF, CTR_LOCATED, CTR_CLUSTERS, CTR_N };

__device__ __forceinline__ void flush_counters(unsigned long long *s_ctr, SeedCounters *ctr) {
    __syncthreads();
    if (threadIdx.x < CTR_N && s_ctr[threadIdx.x]) {
        unsigned long long *dst = threadIdx.x == CTR_SEARCHES ? &ctr->searches
                                  : threadIdx.x == CTR_LF     ? &ctr->lf_steps
                                  : threadIdx.x == CTR_HBM_LF ? &ctr->hbm_lf_steps
                                  : threadIdx.x == CTR_LOCATED ? &ctr->located
                                                               : &ctr->clusters;
        atomicAdd(dst, s_ctr[threadIdx.x]);
    }
}

// ---------------------------------------------------------------------------------------------------
// bwa:find_seeds/6 for a batch of independent subsequences: one thread per search.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) search_kernel(DevIndex ix, const uint8_t *__restrict__ subs,
                                                     const int64_t *__restrict__ offs, int64_t n, int32_t fixed_len,
                                                     int32_t *__restrict__ status, uint32_t *__restrict__ sp,
                                                     uint32_t *__restrict__ ep, int32_t *__restrict__ nm,
                                                     SeedCounters *ctr) {
    __shared__ unsigned long long s_ctr[CTR_N];
    if (threadIdx.x < CTR_N) s_ctr[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long lf = 0, hbm = 0, cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *q;
        int len;
        if (fixed_len > 0) { q = subs + i * fixed_len; len = fixed_len; }
        else { q = subs + offs[i]; len = (int)(offs[i + 1] - offs[i]); }
        SearchRes r;
        if (len <= 0) { r.sp = r.ep = 0; r.n = 0; r.status = BWT_SEARCH_CRASH; r.table = 0; }   // hd([]) badmatch
        else r = search_sub(ix, q, len, true);
        if (status) status[i] = r.status;
        if (sp) sp[i] = r.sp;
        if (ep) ep[i] = r.ep;
        if (nm) nm[i] = r.n;
        cnt++;
        int steps = r.n > 0 ? r.n - 1 : 0;
        lf += steps;
        hbm += r.table ? r.n - r.table : steps;
    }
    // warp-reduce, then one shared atomic per warp
    for (int d = 16; d > 0; d >>= 1) {
        lf += __shfl_down_sync(0xFFFFFFFFu, lf, d);
        hbm += __shfl_down_sync(0xFFFFFFFFu, hbm, d);
        cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_ctr[CTR_SEARCHES], cnt);
        atomicAdd(&s_ctr[CTR_LF], lf);
        atomicAdd(&s_ctr[CTR_HBM_LF], hbm);
    }
    if (ctr) flush_counters(s_ctr, ctr);
}

// `emitted` slots in this block's sub-list (by block, not by read: neighbouring reads stay neighbours in the dense list,
// which emit_kernel's gathers like: by read number emit went 0.71 -> 1.03 ms) -> sparse index of the first one (the region's end when it is full: every
// write is then refused by cluster_slot_ok and the host retries with a larger array)
__device__ __forceinline__ uint32_t cluster_alloc(SeedCounters *ctr, uint32_t read, uint32_t emitted, uint32_t cl_cap) {
    (void)read;
    const uint32_t r = blockIdx.x % SGA_SUBLISTS, cap_r = cl_cap / SGA_SUBLISTS;
    const uint32_t local = atomicAdd(&ctr->sub_n[32 * r], emitted);
    if ((uint64_t)local + emitted > cap_r) atomicExch(&ctr->cluster_overflow, 1u);
    return r * cap_r + min(local, cap_r);
}
__device__ __forceinline__ bool cluster_slot_ok(uint32_t base_out, uint32_t read, uint32_t rank, uint32_t cl_cap) {
    (void)read;
    const uint32_t r = blockIdx.x % SGA_SUBLISTS, cap_r = cl_cap / SGA_SUBLISTS;
    return (uint64_t)(base_out - r * cap_r) + rank < cap_r;
}

// ---------------------------------------------------------------------------------------------------
// sga:skip_Ns/1 (sga.erl:69-76): -1 = no_more_subseqs, else Skipped (the new length is len - Skipped).
// The 'N' that ends a skipped stretch is itself counted as the first of the next eight.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int skip_ns(const uint8_t *q, int len) {
    if (len < BWT_MIN_LEN) return -1;
    int j = 0, skipped = 0, k = len - 1;
    for (;;) {
        if (j == BWT_MIN_LEN) return skipped;
        if (k < 0) return -1;
        if (q[k] == 'N') { skipped += j; j = 1; } else j++;
        k--;
    }
}

struct GroupShared {        // per-group scratch for the task list of one read
    uint16_t *t_end;        // subsequence end (exclusive) inside the read
    uint16_t *t_shift;      // Shift+Skip added to its seed ends (sga.erl:32-33,37-39)
    uint2 *t_res;           // {Sp, count<<4 | N}
    uint32_t *t_pos;        // first slot of its seed ends
};

enum { SGA_DONE = 0, SGA_DEFER = 1 };

// One read, one group of G lanes.  `seeds` holds `cap` uint32 (a power of two) in shared or global memory.
// q: the read's ASCII text, or null when it is a plain (packed) read: then pkr points at its packed words.
template <int G, int MAXT, bool SLOW, typename Tile>
__device__ __forceinline__ int sga_read(const Tile &tile, const DevIndex &ix, const uint8_t *__restrict__ q,
                                        const uint32_t *__restrict__ pkr, int lq,
                                        uint32_t read, uint32_t *seeds, uint32_t cap, const GroupShared &gs,
                                        int32_t *read_status, uint32_t *ncl, uint32_t *cl_base, Cluster *clusters,
                                        uint32_t cl_cap, SlowRead *slow, uint32_t slow_cap, SeedCounters *ctr,
                                        unsigned long long *s_ctr) {
    const int lane = tile.thread_rank();
    if (lq > BWT_MAX_READ_LEN) {
        if (lane == 0) { read_status[read] = BWT_READ_LIMIT; ncl[read] = 0; cl_base[read] = 0; }
        return SGA_DONE;
    }
    // -- task list: where each subsequence ends and what is added to its seed ends
    bool has_n = false;
    if (q) {
        for (int i = lane; i < lq; i += G) has_n |= (q[i] == 'N');
        has_n = tile.any(has_n);
    }
    int nt;
    if (!has_n) {
        nt = lq >= BWT_MIN_LEN ? (lq - BWT_MIN_LEN) / BWT_MIN_LEN + 1 : 0;
    } else {
        nt = 0;
        if (lane == 0) {                                       // sga.erl:18-34, the search-independent part
            int len = lq, shift = 0;
            for (;;) {
                int skip = skip_ns(q, len);
                if (skip < 0) break;
                len -= skip;
                gs.t_end[nt] = (uint16_t)len;
                gs.t_shift[nt] = (uint16_t)(shift + skip);
                nt++;
                len -= BWT_MIN_LEN;
                shift += BWT_MIN_LEN + skip;
            }
        }
        nt = tile.shfl(nt, 0);
        tile.sync();
    }
    // -- searches: lane k takes subsequences k, k+G, ...
    bool crash = false;
    unsigned long long lf = 0, hbm = 0;
    for (int k = lane; k < nt; k += G) {
        int end = has_n ? (int)gs.t_end[k] : lq - BWT_MIN_LEN * k;
        SearchRes r;
        if (q) r = search_sub(ix, q, end, true);
        else r = search_packed(ix, pkr, lq, end, __funnelshift_r(__ldg(pkr + (k >> 1)), __ldg(pkr + (k >> 1) + 1), (k & 1) * 16) & 0x3FFFFFFFu);
        crash |= (r.status == BWT_SEARCH_CRASH);
        uint32_t cnt = (r.status == BWT_SEEDS) ? (r.ep - r.sp + 1) : 0u;
        if (cnt > BWT_MAX_SEEDS_PER_READ) cnt = BWT_MAX_SEEDS_PER_READ + 1;     // flagged below
        gs.t_res[k] = make_uint2(r.sp, (cnt << 4) | (uint32_t)r.n);
        int steps = r.n > 0 ? r.n - 1 : 0;
        lf += steps;
        hbm += r.table ? r.n - r.table : steps;
    }
    crash = tile.any(crash);
    tile.sync();
    // -- slots: exclusive prefix of the seed counts; Qty = searches that produced a list (sga.erl:30-33)
    uint32_t total = 0, qty = 0;
    for (int k0 = 0; k0 < nt; k0 += G) {
        int k = k0 + lane;
        uint32_t c = k < nt ? (gs.t_res[k].y >> 4) : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            uint32_t t = tile.shfl_up(inc, d);
            if (lane >= d) inc += t;
        }
        if (k < nt) gs.t_pos[k] = total + inc - c;
        total += tile.shfl(inc, G - 1);
        qty += __popc(tile.ballot(c > 0));
    }
    if (!SLOW && lane == 0) atomicAdd(&s_ctr[CTR_SEARCHES], (unsigned long long)nt);
    // lane-sum of the LF counters
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) {
        lf += tile.shfl_down(lf, d);
        hbm += tile.shfl_down(hbm, d);
    }
    if (!SLOW && lane == 0) {
        atomicAdd(&s_ctr[CTR_LF], lf);
        atomicAdd(&s_ctr[CTR_HBM_LF], hbm);
    }
    if (crash || total > BWT_MAX_SEEDS_PER_READ) {
        if (lane == 0) {
            read_status[read] = crash ? BWT_READ_CRASH : BWT_READ_LIMIT;
            ncl[read] = 0; cl_base[read] = 0;
        }
        return SGA_DONE;
    }
    if (lane == 0) read_status[read] = BWT_READ_OK;
    if (qty < 2) {                                             // get_similar(0|1,_) -> []  (sga.erl:42-43)
        if (lane == 0) { ncl[read] = 0; cl_base[read] = 0; }
        if (!SLOW && lane == 0) atomicAdd(&s_ctr[CTR_LOCATED], (unsigned long long)total);
        return SGA_DONE;
    }
    if (total > cap) {                                         // does not fit on chip: hand to the slow kernel
        if (lane == 0) {
            uint32_t slot = atomicAdd(&ctr->n_slow, 1u);
            if (slot < slow_cap) { slow[slot].read = read; slow[slot].seeds = total; }
            ncl[read] = 0; cl_base[read] = 0;          // until the slow kernel has done the read
        }
        return SGA_DEFER;
    }
    if (lane == 0) {
        if (SLOW) atomicAdd(&ctr->located, (unsigned long long)total);
        else atomicAdd(&s_ctr[CTR_LOCATED], (unsigned long long)total);
    }
    // -- locate: SA of every row of every interval, plus N and the subsequence's shift (fmi.erl:89-100, sga.erl:37-39)
    uint32_t P = 1;
    while (P < total) P <<= 1;
    if (SLOW) {
        for (int k = 0; k < nt; k++) {
            uint2 tr = gs.t_res[k];
            uint32_t cnt = tr.y >> 4, add = (tr.y & 15u) + (has_n ? (uint32_t)gs.t_shift[k] : (uint32_t)(BWT_MIN_LEN * k));
            uint32_t pos = gs.t_pos[k];
            for (uint32_t t = lane; t < cnt; t += G) seeds[pos + t] = __ldg(ix.sa + (tr.x - 1) + t) + add;
        }
    } else {
        for (int k = lane; k < nt; k += G) {
            uint2 tr = gs.t_res[k];
            uint32_t cnt = tr.y >> 4, add = (tr.y & 15u) + (has_n ? (uint32_t)gs.t_shift[k] : (uint32_t)(BWT_MIN_LEN * k));
            uint32_t pos = gs.t_pos[k];
            for (uint32_t t = 0; t < cnt; t++) seeds[pos + t] = __ldg(ix.sa + (tr.x - 1) + t) + add;
        }
    }
    for (uint32_t i = total + lane; i < P; i += G) seeds[i] = 0xFFFFFFFFu;
    tile.sync();
    // -- lists:sort/1 (sga.erl:45): bitonic network over P slots
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < P; i += G) {
                uint32_t x = i ^ j;
                if (x > i) {
                    uint32_t a = seeds[i], b = seeds[x];
                    bool asc = (i & k) == 0;
                    if ((a > b) == asc) { seeds[i] = b; seeds[x] = a; }
                }
            }
            tile.sync();
        }
    }
    // -- get_similar/7 (sga.erl:46-64): chains with gaps <= TOLERANCE; keep those with >= 2 + Qty div 7
    //    members as {last position, sum of gaps}; the reference's list is descending by position.
    const uint32_t need = 2 + qty / 7;
    uint32_t base_out = 0, emitted = 0;
    for (int pass = 0; pass < 2; pass++) {
        emitted = 0;
        for (uint32_t b0 = 0; b0 < total; b0 += G) {
            int64_t i = (int64_t)total - 1 - (int64_t)(b0 + lane);
            bool em = false;
            uint32_t S = 0, D = 0;
            if (i >= 0) {
                uint32_t si = seeds[i];
                bool is_end = (i == (int64_t)total - 1) || (seeds[i + 1] - si > BWT_TOLERANCE);
                if (is_end) {
                    int64_t k = i;
                    while (k > 0 && seeds[k] - seeds[k - 1] <= BWT_TOLERANCE) k--;
                    if ((uint32_t)(i - k + 1) >= need) { em = true; S = si; D = si - seeds[k]; }
                }
            }
            uint32_t bal = tile.ballot(em);
            if (pass == 1 && em) {
                uint32_t rank = emitted + __popc(bal & ((1u << lane) - 1u));
                if (cluster_slot_ok(base_out, read, rank, cl_cap)) { // on overflow the host retries with the needed capacity
                    Cluster c;
                    c.read = read; c.s = S; c.d = (int32_t)D; c.rank = rank;
                    clusters[base_out + rank] = c;
                }
            }
            emitted += __popc(bal);
        }
        if (pass == 0) {
            if (emitted == 0) break;
            if (lane == 0) {
                base_out = cluster_alloc(ctr, read, emitted, cl_cap);
            }
            base_out = tile.shfl(base_out, 0);
        }
    }
    if (lane == 0) {
        ncl[read] = emitted;
        cl_base[read] = base_out;
        if (!SLOW) atomicAdd(&s_ctr[CTR_CLUSTERS], (unsigned long long)emitted);
        else atomicAdd(&ctr->clusters, (unsigned long long)emitted);
    }
    tile.sync();
    return SGA_DONE;
}

// ---------------------------------------------------------------------------------------------------
// Fast path of sga:sga/6 for the common read: plain ACGT, at most 8*G seed ends.
//   * the read is packed once into 2-bit codes, REVERSED (base lq-1-i at bits 2i), in shared memory, so the k-mer
//     table index of the subsequence that ends 8k bases before the read's end is one funnel shift at bit 16k
//   * seed ends are located into shared memory by all lanes (balanced over the rows of all intervals), then sorted
//     in REGISTERS (bitonic network, EPL consecutive elements per lane, shuffles across lanes) and voted there with
//     a segmented scan -- no shared-memory sort, no second voting pass
// ---------------------------------------------------------------------------------------------------
enum { SGA_FALLBACK = 2 };

template <int G, int EPL, typename Tile>
__device__ __forceinline__ void reg_bitonic(const Tile &tile, uint32_t (&v)[EPL], const int lane) {
    constexpr int P = G * EPL;
#pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= EPL) {                                      // partner in another lane, same register
                const int lj = j / EPL;
                const bool up = ((lane * EPL) & k) == 0;
                const bool keep_min = ((lane & lj) == 0) == up;
#pragma unroll
                for (int r = 0; r < EPL; r++) {
                    const uint32_t o = tile.shfl_xor(v[r], lj);
                    v[r] = keep_min ? min(v[r], o) : max(v[r], o);
                }
            } else {                                             // partner in this lane
#pragma unroll
                for (int r = 0; r < EPL; r++) {
                    if ((r & j) == 0) {
                        const bool up = (k >= EPL) ? (((lane * EPL) & k) == 0) : ((r & k) == 0);
                        const uint32_t a = v[r], b = v[r | j];
                        const uint32_t lo = min(a, b), hi = max(a, b);
                        v[r] = up ? lo : hi;
                        v[r | j] = up ? hi : lo;
                    }
                }
            }
        }
    }
}

// sort + get_similar/7 (sga.erl:45-64) on `total` seed ends in shared memory; emits the clusters
template <int G, int EPL, typename Tile>
__device__ __forceinline__ void sort_vote_regs(const Tile &tile, const uint32_t *seeds, const uint32_t total,
                                               const uint32_t need, const uint32_t read, uint32_t *ncl, uint32_t *cl_base,
                                               Cluster *clusters, const uint32_t cl_cap, SeedCounters *ctr,
                                               uint32_t (&acc)[CTR_N]) {
    const int lane = tile.thread_rank();
    uint32_t v[EPL];
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const uint32_t i = (uint32_t)lane * EPL + r;
        v[r] = i < total ? seeds[i] : 0xFFFFFFFFu;
    }
    reg_bitonic<G, EPL>(tile, v, lane);
    // chain breaks: element i opens a chain when it is the first or lies more than TOLERANCE after its predecessor
    const uint32_t prev_last = tile.shfl_up(v[EPL - 1], 1);
    const uint32_t next_first = tile.shfl_down(v[0], 1);
    bool f = false;                                              // lane summary: start of the chain running at its end
    uint32_t sv = 0, si = 0;
    uint32_t brk = 0;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const uint32_t i = (uint32_t)lane * EPL + r;
        const uint32_t pv = r ? v[r - 1] : prev_last;
        const bool b = (i == 0) || (v[r] - pv > BWT_TOLERANCE);
        brk |= (uint32_t)b << r;
        if (b) { f = true; sv = v[r]; si = i; }
    }
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {                            // inclusive scan: nearest chain start at or before the lane's end
        const bool f2 = tile.shfl_up(f, d);
        const uint32_t sv2 = tile.shfl_up(sv, d), si2 = tile.shfl_up(si, d);
        if (lane >= d && !f) { f = f2; sv = sv2; si = si2; }
    }
    uint32_t cur_sv = tile.shfl_up(sv, 1), cur_si = tile.shfl_up(si, 1);   // chain running into this lane (lane 0: element 0 breaks)
    uint32_t em = 0, dist[EPL];
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const uint32_t i = (uint32_t)lane * EPL + r;
        if ((brk >> r) & 1u) { cur_sv = v[r]; cur_si = i; }
        const uint32_t nv = (r + 1 < EPL) ? v[(r + 1) % EPL] : next_first;
        const bool is_end = i < total && (i + 1 == total || nv - v[r] > BWT_TOLERANCE);
        dist[r] = v[r] - cur_sv;
        if (is_end && i - cur_si + 1 >= need) em |= 1u << r;
    }
    // the reference's list is descending by position: rank = number of emitted elements with a higher index
    const uint32_t c = __popc(em);
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const uint32_t t = tile.shfl_down(inc, d);
        if (lane + d < G) inc += t;
    }
    const uint32_t emitted = tile.shfl(inc, 0);
    uint32_t base_out = 0;
    if (emitted) {
        if (lane == 0) {
            base_out = cluster_alloc(ctr, read, emitted, cl_cap);
        }
        base_out = tile.shfl(base_out, 0);
        uint32_t rank = inc - c;                                 // emitted in higher lanes
#pragma unroll
        for (int r = EPL - 1; r >= 0; r--) {
            if ((em >> r) & 1u) {
                if (cluster_slot_ok(base_out, read, rank, cl_cap)) {
                    Cluster cc;
                    cc.read = read; cc.s = v[r]; cc.d = (int32_t)dist[r]; cc.rank = rank;
                    clusters[base_out + rank] = cc;
                }
                rank++;
            }
        }
    }
    if (lane == 0) {
        ncl[read] = emitted;
        cl_base[read] = base_out;
        acc[CTR_CLUSTERS] += emitted;
    }
}

// A miss of a plain load makes the L2 fetch the whole 128-byte line (ncu on scripts/probes/gather_probe.cu: 127-137 B of
// DRAM per random 16/32-byte access); with the L2::64B prefetch-size qualifier it fetches 64 bytes (69 B per access).
__device__ __forceinline__ uint4 ldg_l2_64(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// One plain read (packed words pkr, lq bases).  Per subsequence ONE lookup: the fused depth-15 entry (interval + its
// first 14 SA values in one 64-byte line) when the subsequence has 15 characters or more, else the interval table of
// exactly its depth (+ the suffix-array run).  The lane that looked a subsequence up writes its seed ends.
template <int G, int MAXT, typename Tile>
__device__ __forceinline__ int sga_fast_read(const Tile &tile, const DevIndex &ix, const uint32_t *__restrict__ pkr, const int lq,
                                             const uint32_t read, uint32_t *seeds,
                                             int32_t *read_status, uint32_t *ncl, uint32_t *cl_base, Cluster *clusters,
                                             const uint32_t cl_cap, SeedCounters *ctr, uint32_t (&acc)[CTR_N], const bool hint64) {
    constexpr int EPLMAX = G == 8 ? 16 : 8;                       // seed ends per lane the register sort takes: 128 per read either way
    const int lane = tile.thread_rank();
    const int nt = lq >= BWT_MIN_LEN ? (lq - BWT_MIN_LEN) / BWT_MIN_LEN + 1 : 0;
    uint32_t total = 0, qty = 0, lf = 0, hbm = 0;
    bool fb = false;
    for (int k0 = 0; k0 < nt; k0 += G) {
        const int k = k0 + lane;
        uint32_t cnt = 0, sp = 0, add = 0;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        const uint4 *e = nullptr;
        if (k < nt) {
            const int end = lq - BWT_MIN_LEN * k;                 // subsequence k ends 8k bases before the end of the read
            const uint32_t idx = __funnelshift_r(__ldg(pkr + (k >> 1)), __ldg(pkr + (k >> 1) + 1), (k & 1) * 16) & 0x3FFFFFFFu;
            if (end >= BWT_MAX_LEN && ix.fused) {
                e = ix.fused + 4 * (size_t)idx;
                if (hint64) { a = ldg_l2_64(e); b = ldg_l2_64(e + 1); }   // one 64-byte entry: do not drag the 128-byte line in
                else { a = __ldg(e); b = __ldg(e + 1); }
                const uint32_t ne = a.y >> 28, w = a.y & KMER_W_MASK;
                if (ne == 0) fb = true;                            // width beyond 28 bits: the general path walks it
                else {
                    if (ne == BWT_MAX_LEN) cnt = w > BWT_MAX_RANGE ? 0u : w + 1;     // bwa.erl:31-32
                    else if (ne >= BWT_MIN_LEN) cnt = w + 1;                          // not_found stop: no width check
                    sp = a.x; add = ne + (uint32_t)(BWT_MIN_LEN * k);
                    lf += ne - 1;
                }
            } else {
                const SearchRes r = search_packed(ix, pkr, lq, end, idx);
                if (r.status == BWT_SEEDS) cnt = r.ep - r.sp + 1;
                sp = r.sp; add = (uint32_t)r.n + (uint32_t)(BWT_MIN_LEN * k);
                const uint32_t steps = r.n > 0 ? r.n - 1 : 0;
                lf += steps;
                hbm += r.table ? r.n - r.table : steps;
            }
            if (cnt > BWT_MAX_SEEDS_PER_READ) cnt = BWT_MAX_SEEDS_PER_READ + 1;
        }
        // slots: exclusive prefix of the seed counts; Qty = searches that produced a list (sga.erl:30-33)
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const uint32_t t = tile.shfl_up(inc, d);
            if (lane >= d) inc += t;
        }
        const uint32_t pos = total + inc - cnt;
        total += tile.shfl(inc, G - 1);
        qty += __popc(tile.ballot(cnt > 0));
        // seed ends = SA + N + the subsequence's shift (fmi.erl:89-100, sga.erl:37-39); beyond 8G the read leaves this path
        if (cnt && pos + cnt <= (uint32_t)(EPLMAX * G)) {
            uint32_t t0 = 0;
            if (e) {
                const uint32_t v[6] = {a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int t = 0; t < 6; t++) if ((uint32_t)t < cnt) seeds[pos + t] = v[t] + add;
                if (cnt > 6) {
                    const uint4 c = hint64 ? ldg_l2_64(e + 2) : __ldg(e + 2), d = hint64 ? ldg_l2_64(e + 3) : __ldg(e + 3);
                    const uint32_t u[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int t = 0; t < 8; t++) if ((uint32_t)(t + 6) < cnt) seeds[pos + 6 + t] = u[t] + add;
                }
                t0 = 14;
            }
            for (uint32_t t = t0; t < cnt; t++) seeds[pos + t] = __ldg(ix.sa + (sp - 1) + t) + add;
        }
    }
    if (tile.any(fb) || total > (uint32_t)(EPLMAX * G)) return SGA_FALLBACK;   // the general path redoes (and counts) the searches (rare)
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) {
        lf += tile.shfl_down(lf, d);
        hbm += tile.shfl_down(hbm, d);
    }
    if (lane == 0) {
        acc[CTR_SEARCHES] += (uint32_t)nt; acc[CTR_LF] += lf; acc[CTR_HBM_LF] += hbm; acc[CTR_LOCATED] += total;
        read_status[read] = BWT_READ_OK;
    }
    if (qty < 2) {                                             // get_similar(0|1,_) -> []  (sga.erl:42-43)
        if (lane == 0) { ncl[read] = 0; cl_base[read] = 0; }
        return SGA_DONE;
    }
    tile.sync();
    const uint32_t need = 2 + qty / 7;
    if (total <= (uint32_t)G) sort_vote_regs<G, 1>(tile, seeds, total, need, read, ncl, cl_base, clusters, cl_cap, ctr, acc);
    else if (total <= (uint32_t)(2 * G)) sort_vote_regs<G, 2>(tile, seeds, total, need, read, ncl, cl_base, clusters, cl_cap, ctr, acc);
    else if (total <= (uint32_t)(4 * G)) sort_vote_regs<G, 4>(tile, seeds, total, need, read, ncl, cl_base, clusters, cl_cap, ctr, acc);
    else if (EPLMAX == 8 || total <= (uint32_t)(8 * G)) sort_vote_regs<G, 8>(tile, seeds, total, need, read, ncl, cl_base, clusters, cl_cap, ctr, acc);
    else sort_vote_regs<G, EPLMAX>(tile, seeds, total, need, read, ncl, cl_base, clusters, cl_cap, ctr, acc);
    tile.sync();
    return SGA_DONE;
}

// The fused entries of the group's NEXT read, asked into L2 while the current read is looked up, sorted and voted on:
// a group otherwise has loads in flight only while it waits for its own 13 entries (~1 us) and none during the ~40 %
// of its time it spends on the seed ends, and random access on this part is bound by the misses in flight, not by
// DRAM bytes (scripts/probes/gather_probe.cu: 35-47 G random accesses/s whether a miss moves 64 or 128 bytes).
// mode 1: prefetch.global.L2 (a 128-byte line), mode 2: cp.async.bulk.prefetch.L2 of exactly the 64-byte entry (a
// warp-uniform instruction: the compiler loops over the lanes), mode 3: a discarded ld.global.nc.L2::64B.
template <int G>
__device__ __forceinline__ void sga_prefetch_next(const DevIndex &ix, const DevReads &R, const int64_t nxt, const int lane, const int mode) {
    const int lq = (int)__ldg(R.len + nxt);
    const uint32_t *pkr = R.pk + __ldg(R.woff + nxt);
    const int nt = lq >= BWT_MIN_LEN ? (lq - BWT_MIN_LEN) / BWT_MIN_LEN + 1 : 0;
    for (int k = lane; k < nt; k += G) {
        if (lq - BWT_MIN_LEN * k < BWT_MAX_LEN) continue;              // the short front subsequence uses the small tables
        const uint32_t idx = __funnelshift_r(__ldg(pkr + (k >> 1)), __ldg(pkr + (k >> 1) + 1), (k & 1) * 16) & 0x3FFFFFFFu;
        const uint4 *e = ix.fused + 4 * (size_t)idx;
        if (mode == 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(e));
        else if (mode == 3) asm volatile("{ .reg .b32 t; ld.global.nc.L2::64B.b32 t, [%0]; }" :: "l"(e));   // a load nobody waits for
        else asm volatile("cp.async.bulk.prefetch.L2.global [%0], 64;" :: "l"(e) : "memory");
    }
}

template <int G, int MAXT>
__global__ void __launch_bounds__(SGA_BLOCK, G == 16 ? SGA_MIN_BLOCKS : 8) sga_kernel(DevIndex ix, DevReads R, int64_t n_reads,
                                                        int32_t *read_status, uint32_t *ncl, uint32_t *cl_base,
                                                        Cluster *clusters, uint32_t cl_cap, SlowRead *slow,
                                                        uint32_t slow_cap, SeedCounters *ctr, int prefetch) {
    constexpr int GPB = SGA_BLOCK / G;
    constexpr int CAP = 32 * G;
    __shared__ uint32_t s_seeds[GPB][CAP];
    __shared__ uint16_t s_end[GPB][MAXT], s_shift[GPB][MAXT];
    __shared__ uint2 s_res[GPB][MAXT];
    __shared__ uint32_t s_pos[GPB][MAXT];
    __shared__ unsigned long long s_ctr[CTR_N];
    if (threadIdx.x < CTR_N) s_ctr[threadIdx.x] = 0;
    __syncthreads();
    auto tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int g = threadIdx.x / G;
    GroupShared gs{s_end[g], s_shift[g], s_res[g], s_pos[g]};
    uint32_t acc[CTR_N] = {0u, 0u, 0u, 0u, 0u};                   // lane 0 of the group: work counters of its fast-path reads
    for (int64_t read = (int64_t)blockIdx.x * GPB + g; read < n_reads; read += (int64_t)gridDim.x * GPB) {
        int lq = (int)__ldg(R.len + read);
        const bool plain = !read_is_exc(R, (uint32_t)read);
        const uint8_t *q = nullptr;
        const uint32_t *pkr = nullptr;
        if ((prefetch & 7) && ix.fused && read + (int64_t)gridDim.x * GPB < n_reads)
            sga_prefetch_next<G>(ix, R, read + (int64_t)gridDim.x * GPB, (int)tile.thread_rank(), prefetch & 7);
        if (plain) pkr = R.pk + __ldg(R.woff + read);
        else q = read_ascii(R, (uint32_t)read, lq);
        // a read with more subsequences than MAXT cannot run in this instantiation: the launcher picks
        // G/MAXT from the longest read of the batch, so this only guards against misuse
        if (lq > BWT_MIN_LEN * (MAXT - 1) + BWT_MAX_LEN && MAXT < SGA_MAX_TASKS) {
            if (tile.thread_rank() == 0) { read_status[read] = BWT_READ_LIMIT; ncl[read] = 0; cl_base[read] = 0; }
            continue;
        }
        int how = SGA_FALLBACK;
        if (plain && lq <= BWT_MIN_LEN * (MAXT - 1) + BWT_MAX_LEN)
            how = sga_fast_read<G, MAXT>(tile, ix, pkr, lq, (uint32_t)read, s_seeds[g], read_status, ncl, cl_base,
                                         clusters, cl_cap, ctr, acc, (prefetch & 8) != 0);
        if (how == SGA_FALLBACK)                                  // 'N' / other bytes, or more seed ends than registers hold
            sga_read<G, MAXT, false>(tile, ix, q, pkr, lq, (uint32_t)read, s_seeds[g], CAP, gs, read_status, ncl, cl_base,
                                     clusters, cl_cap, slow, slow_cap, ctr, s_ctr);
        tile.sync();
    }
    if (tile.thread_rank() == 0) {
#pragma unroll
        for (int k = 0; k < CTR_N; k++)
            if (acc[k]) atomicAdd(&s_ctr[k], (unsigned long long)acc[k]);
    }
    flush_counters(s_ctr, ctr);
}

// reads whose seed ends overflow the shared-memory budget: one warp per read, seeds in an HBM scratch handed out by a
// device-side cursor (scratch_cap words; a read that does not fit raises scratch_overflow and the host retries larger)
__global__ void __launch_bounds__(SGA_BLOCK) sga_slow_kernel(DevIndex ix, DevReads R,
                                                             const SlowRead *__restrict__ slow, uint32_t slow_cap,
                                                             uint32_t *scratch, unsigned long long scratch_cap,
                                                             int32_t *read_status, uint32_t *ncl,
                                                             uint32_t *cl_base, Cluster *clusters, uint32_t cl_cap,
                                                             SeedCounters *ctr) {
    constexpr int G = 32, GPB = SGA_BLOCK / G, MAXT = SGA_MAX_TASKS;
    __shared__ uint16_t s_end[GPB][MAXT], s_shift[GPB][MAXT];
    __shared__ uint2 s_res[GPB][MAXT];
    __shared__ uint32_t s_pos[GPB][MAXT];
    auto tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int g = threadIdx.x / G;
    GroupShared gs{s_end[g], s_shift[g], s_res[g], s_pos[g]};
    const uint32_t n_slow = min(ctr->n_slow, slow_cap);
    for (uint32_t w = blockIdx.x * GPB + g; w < n_slow; w += gridDim.x * GPB) {
        const uint32_t read = slow[w].read;
        unsigned long long p = 1;
        while (p < slow[w].seeds) p <<= 1;
        unsigned long long base = 0;
        if (tile.thread_rank() == 0) base = atomicAdd(&ctr->scratch_used, p);
        base = tile.shfl(base, 0);
        if (base + p > scratch_cap) {
            if (tile.thread_rank() == 0) { atomicExch(&ctr->scratch_overflow, 1u); ncl[read] = 0; cl_base[read] = 0; }
            continue;
        }
        int lq = (int)R.len[read];
        const uint8_t *q = nullptr;
        const uint32_t *pkr = nullptr;
        if (!read_is_exc(R, read)) pkr = R.pk + R.woff[read];
        else q = read_ascii(R, read, lq);
        sga_read<G, MAXT, true>(tile, ix, q, pkr, lq, read, scratch + base, (uint32_t)p, gs, read_status, ncl,
                                cl_base, clusters, cl_cap, nullptr, 0, ctr, nullptr);
        tile.sync();
    }
}

}  // namespace

namespace {
__global__ void cluster_offsets_kernel(SeedCounters *ctr, uint32_t cap_r, uint32_t slow_cap) {     // one warp
    uint32_t sum = 0, dense = 0;
    for (int r = 0; r < SGA_SUBLISTS; r++) {
        const uint32_t c = ctr->sub_n[32 * r];
        if (threadIdx.x == 0) ctr->dense_base[r] = dense;
        sum += c; dense += min(c, cap_r);
    }
    if (threadIdx.x == 0) {
        // a capacity overflowed (cluster sub-list, deferred-read list, slow scratch): per-read results are incomplete, so the
        // kernels behind this one get nothing to do and the host grows the capacity and runs the batch again
        const uint32_t ab = (ctr->cluster_overflow | ctr->scratch_overflow | (ctr->n_slow > slow_cap ? 1u : 0u)) ? 1u : 0u;
        ctr->n_clusters = sum;
        ctr->abort = ab;
        ctr->n_pairs = ab ? 0u : sum;
    }
}
__global__ void cluster_compact_kernel(const Cluster *__restrict__ sparse, Cluster *__restrict__ dense, const SeedCounters *ctr,
                                       uint32_t cap_r) {
    const uint32_t r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < min(ctr->sub_n[32 * r], cap_r)) dense[ctr->dense_base[r] + i] = sparse[(size_t)r * cap_r + i];
}
__global__ void cluster_rebase_kernel(uint32_t *__restrict__ cl_base, const uint32_t *__restrict__ ncl, int64_t n,
                                      const SeedCounters *ctr, uint32_t cap_r) {
    const int64_t read = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (read >= n || ncl[read] == 0) return;
    const uint32_t b = cl_base[read], r = min(b / cap_r, (uint32_t)SGA_SUBLISTS - 1);
    cl_base[read] = ctr->dense_base[r] + (b - r * cap_r);
}
}  // namespace

int launch_sga_finish(const Cluster *d_sparse, Cluster *d_dense, uint32_t cl_cap, uint32_t *d_cl_base, const uint32_t *d_ncl,
                      int64_t n_reads, uint32_t slow_cap, SeedCounters *d_ctr, cudaStream_t stream, uint32_t *launches) {
    const uint32_t cap_r = cl_cap / SGA_SUBLISTS;
    cluster_offsets_kernel<<<1, 32, 0, stream>>>(d_ctr, cap_r, slow_cap);
    cluster_compact_kernel<<<dim3((cap_r + 255) / 256, SGA_SUBLISTS), 256, 0, stream>>>(d_sparse, d_dense, d_ctr, cap_r);
    if (n_reads > 0) cluster_rebase_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, stream>>>(d_cl_base, d_ncl, n_reads, d_ctr, cap_r);
    if (launches) *launches += 3;
    K_CHECK(stream, "cluster_offsets / compact / rebase kernels");
    return BWT_OK;
}

int launch_search(const DevIndex &ix, const uint8_t *d_subs, const int64_t *d_offs, int64_t n, int32_t fixed_len,
                  int32_t *d_status, uint32_t *d_sp, uint32_t *d_ep, int32_t *d_n, SeedCounters *d_ctr,
                  cudaStream_t stream) {
    if (n <= 0) return BWT_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;         // grid-stride beyond 64 CTAs per SM
    search_kernel<<<(unsigned)blocks, 256, 0, stream>>>(ix, d_subs, d_offs, n, fixed_len, d_status, d_sp, d_ep, d_n, d_ctr);
    CU_TRY(cudaGetLastError());
    return BWT_OK;
}

int launch_sga(const DevIndex &ix, const DevReads &R, int64_t n_reads, int max_read_len,
               int32_t *d_read_status, uint32_t *d_ncl, uint32_t *d_cl_base, Cluster *d_clusters, uint32_t cl_cap,
               SlowRead *d_slow, uint32_t slow_cap, SeedCounters *d_ctr, cudaStream_t stream) {
    if (n_reads <= 0) return BWT_OK;
    {   // L1 / shared-memory split of the seeding kernels: BWT_SGA_CARVEOUT = percent of shared memory (default: driver's choice)
        static int carve = -2;
        if (carve == -2) { const char *e = getenv("BWT_SGA_CARVEOUT"); carve = e ? atoi(e) : -1; }
        if (carve >= 0) {
            cudaFuncSetAttribute(sga_kernel<16, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(sga_kernel<32, SGA_MAX_TASKS>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        }
    }
    // BWT_SGA_PREFETCH (default 8): +8 = fused entries loaded with the L2::64B qualifier (DRAM read 15.8 -> 8.8 GB per 10 M reads, time +-2 %);
    // low bits: 0 off (default: measured 5-13 % slower with any of the modes), 1 prefetch.global.L2, 2 64-byte bulk prefetch, 3 discarded load
    static int prefetch = -1;
    if (prefetch < 0) { const char *e = getenv("BWT_SGA_PREFETCH"); prefetch = e && *e ? atoi(e) : 8; }
    // 16 lanes per read while every read has at most 16 subsequences (<= 135 bases), else a full warp
    if (max_read_len <= BWT_MIN_LEN * 15 + BWT_MAX_LEN) {
        int64_t blocks = (n_reads + 7) / 8;
        if (blocks > 148 * 256) blocks = 148 * 256;
        sga_kernel<16, 16><<<(unsigned)blocks, SGA_BLOCK, 0, stream>>>(ix, R, n_reads, d_read_status, d_ncl,
                                                                     d_cl_base, d_clusters, cl_cap, d_slow, slow_cap, d_ctr, prefetch);
    } else {
        int64_t blocks = (n_reads + 3) / 4;
        if (blocks > 148 * 256) blocks = 148 * 256;
        sga_kernel<32, SGA_MAX_TASKS><<<(unsigned)blocks, SGA_BLOCK, 0, stream>>>(ix, R, n_reads, d_read_status,
                                                                               d_ncl, d_cl_base, d_clusters, cl_cap, d_slow,
                                                                               slow_cap, d_ctr, prefetch);
    }
    K_CHECK(stream, "sga_kernel");
    return BWT_OK;
}

// always launched (the list length is read on the device): a few blocks that leave at once when no read was deferred
int launch_sga_slow(const DevIndex &ix, const DevReads &R, const SlowRead *d_slow, uint32_t slow_cap,
                    uint32_t *d_scratch, uint64_t scratch_cap, int32_t *d_read_status,
                    uint32_t *d_ncl, uint32_t *d_cl_base, Cluster *d_clusters, uint32_t cl_cap, SeedCounters *d_ctr,
                    cudaStream_t stream) {
    sga_slow_kernel<<<148, SGA_BLOCK, 0, stream>>>(ix, R, d_slow, slow_cap, d_scratch, (unsigned long long)scratch_cap,
                                                   d_read_status, d_ncl, d_cl_base, d_clusters, cl_cap, d_ctr);
    K_CHECK(stream, "sga_slow_kernel");
    return BWT_OK;
}
