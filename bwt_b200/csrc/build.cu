// Index builder on the device: fm_index:make_index/4 (fm_index.erl:127-141) -> bwt:fm/1 (bwt.erl:84-98).
//
// dsa:dsa/1 (dsa.erl:15-48) sorts the suffixes of Seq++"$" under byte order with '$' below ACGT; every correct
// suffix array of that string is the same array, so the construction is free:
//   round 0 : radix sort of all suffixes by their first 21 symbols (3 bits each, '$' = 0 so that suffixes that
//             reach the end are ordered exactly)                                     -- cub::DeviceRadixSort
//   round k : prefix doubling restricted to the still-tied groups: sort (group, rank[i+h]) and split
// CUB's radix sort is the only library primitive used, and only here in the offline builder.
// Then L[e] = Seq[SA[e]-1] ('$' for SA[e] = 0), pointers from the symbol counts (fmi.erl:66-84).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <string.h>
#include <vector>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t sym3(uint8_t c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : 4; }

// key of suffix i: 21 symbols, 3 bits each, most significant first; past the end ('$' and beyond) = 0
__global__ void key0_kernel(const uint8_t *__restrict__ text, uint64_t len, uint64_t n, uint64_t *__restrict__ keys,
                            uint32_t *__restrict__ idx) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
#pragma unroll
    for (int t = 0; t < 21; t++) {
        uint64_t p = i + t;
        k = (k << 3) | (p < len ? sym3(text[p]) : 0ull);
    }
    keys[i] = k;
    idx[i] = (uint32_t)i;
}

// after a sort: head flag of every run of equal keys, within [0,m)
__global__ void heads_kernel(const uint64_t *__restrict__ keys, uint64_t m, uint32_t *__restrict__ head) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    head[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1u : 0u;
}

// group head position of every slot = running maximum of (head ? p : 0): done per block with a serial fix-up
// table of block-leading heads (simple two-level scheme)
__global__ void block_last_head_kernel(const uint32_t *__restrict__ head, uint64_t m, uint32_t *__restrict__ blk_last) {
    // one thread per 1024-slot block: position of the last head inside the block, or 0xFFFFFFFF
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = b * 1024;
    if (lo >= m) return;
    uint64_t hi = lo + 1024 < m ? lo + 1024 : m;
    uint32_t last = 0xFFFFFFFFu;
    for (uint64_t p = lo; p < hi; p++) if (head[p]) last = (uint32_t)p;
    blk_last[b] = last;
}

__global__ void carry_heads_kernel(uint32_t *blk_last, uint64_t nb) {
    // single thread: turn "last head inside block" into "last head at or before the end of the block"
    if (blockIdx.x || threadIdx.x) return;
    uint32_t cur = 0;
    for (uint64_t b = 0; b < nb; b++) {
        if (blk_last[b] != 0xFFFFFFFFu) cur = blk_last[b];
        blk_last[b] = cur;
    }
}

// grp[p] = position of the head of p's group; flags slots of groups with more than one member
__global__ void groups_kernel(const uint32_t *__restrict__ head, uint64_t m, const uint32_t *__restrict__ blk_last,
                              uint32_t *__restrict__ grp, uint32_t *__restrict__ tied) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = b * 1024;
    if (lo >= m) return;
    uint64_t hi = lo + 1024 < m ? lo + 1024 : m;
    uint32_t cur = b ? blk_last[b - 1] : 0u;
    for (uint64_t p = lo; p < hi; p++) {
        if (head[p]) cur = (uint32_t)p;
        grp[p] = cur;
        bool single = head[p] && (p + 1 == m || head[p + 1]);
        tied[p] = single ? 0u : 1u;
    }
}

// round 0 bookkeeping: sa[p] = idx[p]; rank[idx[p]] = grp[p]
__global__ void commit0_kernel(const uint32_t *__restrict__ idx, const uint32_t *__restrict__ grp, uint64_t n,
                               uint32_t *__restrict__ sa, uint32_t *__restrict__ rank) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t i = idx[p];
    sa[p] = i;
    rank[i] = grp[p];
}

// compaction of the tied slots: pos[k] = k-th tied slot (ascending)
__global__ void scatter_tied_kernel(const uint32_t *__restrict__ tied, const uint32_t *__restrict__ off, uint64_t m,
                                    const uint32_t *__restrict__ src_pos, uint32_t *__restrict__ pos) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    if (tied[p]) pos[off[p]] = src_pos ? src_pos[p] : (uint32_t)p;
}

// keys of a doubling round over the tied slots: (group head, rank[i+h]+1 or 0 past the end)
__global__ void keyh_kernel(const uint32_t *__restrict__ pos, uint64_t m, const uint32_t *__restrict__ sa,
                            const uint32_t *__restrict__ rank, uint64_t n, uint64_t h, uint64_t *__restrict__ keys,
                            uint32_t *__restrict__ idx) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t i = sa[pos[k]];
    uint64_t j = (uint64_t)i + h;
    uint64_t second = j < n ? (uint64_t)rank[j] + 1 : 0ull;
    keys[k] = ((uint64_t)rank[i] << 32) | second;
    idx[k] = i;
}

// after sorting a doubling round: slot pos[k] receives idx[k]; its new group head is pos[grp_local[k]]
__global__ void commith_kernel(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ idx,
                               const uint32_t *__restrict__ grp_local, uint64_t m, uint32_t *__restrict__ sa,
                               uint32_t *__restrict__ rank) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t i = idx[k];
    sa[pos[k]] = i;
    rank[i] = pos[grp_local[k]];
}

__global__ void bwt_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ sa, uint64_t n,
                           uint8_t *__restrict__ L) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t i = sa[p];
    L[p] = i ? text[i - 1] : (uint8_t)'$';
}

__global__ void count_kernel(const uint8_t *__restrict__ text, uint64_t len, unsigned long long *__restrict__ cnt) {
    __shared__ unsigned int s[4];
    if (threadIdx.x < 4) s[threadIdx.x] = 0;
    __syncthreads();
    unsigned int c[4] = {0, 0, 0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
        int k = bwt_code(text[i]);
        if (k >= 0) c[k]++;
    }
    for (int k = 0; k < 4; k++) atomicAdd(&s[k], c[k]);
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(&cnt[threadIdx.x], (unsigned long long)s[threadIdx.x]);
}

// IUPAC -> ACGT as fm_index.erl:131-135; flags 'N' and unknown symbols
__global__ void map_acgt_kernel(const uint8_t *__restrict__ in, uint64_t len, uint8_t *__restrict__ out,
                                uint32_t *__restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    uint8_t c = in[i], o;
    switch (c) {
    case 'A': case 'R': case 'M': case 'W': case 'V': o = 'A'; break;
    case 'C': case 'B': case 'Y': case 'S': o = 'C'; break;
    case 'G': case 'D': o = 'G'; break;
    case 'T': case 'K': o = 'T'; break;
    case 'N': o = 'A'; atomicAdd(flags, 1u); break;
    default: o = 'A'; atomicAdd(flags + 1, 1u); break;
    }
    out[i] = o;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {           // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// base i of the synthetic reference (seed): a pure function of the position, so that a partition of it (or a read
// drawn from it) can be produced anywhere without holding the whole text
__device__ __forceinline__ uint8_t synth_base(uint64_t seed, uint64_t i) {
    const uint64_t z = mix64(seed + 0x9E3779B97F4A7C15ull * (i / 32 + 1));   // splitmix64 of the 32-base word index
    return (uint8_t)"ACGT"[(z >> (2 * (i & 31))) & 3];
}

__global__ void synth_kernel(uint8_t *__restrict__ out, uint64_t start, uint64_t len, uint64_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) out[i] = synth_base(seed, start + i);
}

// one thread per base: read r starts at hash(seed, r) mod (ref_len - len + 1); base p of read r is substituted
// (by one of the three other bases of A,C,G,T) with probability thresh / 2^32
__global__ void sample_reads_kernel(const uint8_t *__restrict__ ref, uint64_t ref_len, int64_t n, int32_t len,
                                    uint64_t seed, uint32_t thresh, uint8_t *__restrict__ out,
                                    int64_t *__restrict__ starts) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * len) return;
    int64_t r = t / len;
    int32_t p = (int32_t)(t - r * len);
    uint64_t start = mix64(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(r + 1))) % (ref_len - (uint64_t)len + 1);
    if (p == 0 && starts) starts[r] = (int64_t)start;
    uint8_t c = ref[start + p];
    uint64_t h = mix64((seed + 0x51ED270B1ull) ^ (0xD6E8FEB86659FD93ull * (uint64_t)(t + 1)));
    if ((uint32_t)h < thresh) {
        int k = bwt_code(c);
        if (k >= 0) c = (uint8_t)"ACGT"[(k + 1 + (int)((h >> 32) % 3)) & 3];
    }
    out[t] = c;
}

// SURVEY.md 8(d) edit model, one thread per read: each of the `len` reference bases drawn for read r is edited with
// probability thresh / 2^32 -- substituted, followed by one inserted base (16-bit draw < ins_cut), or deleted
// (< del_cut) -- so the read comes out len + #insertions - #deletions long; afterwards every base of the read turns
// into 'N' with probability n_thresh / 2^32.  out == nullptr: only the length is computed (pass 1).
// ref == nullptr: the bases come from the synthetic reference (synth_seed).
__global__ void sample_edits_kernel(const uint8_t *__restrict__ ref, uint64_t synth_seed, uint64_t ref_len, int64_t n,
                                    int32_t len, uint64_t seed, uint32_t thresh, uint32_t ins_cut, uint32_t del_cut,
                                    uint32_t n_thresh, const int64_t *__restrict__ offs, uint8_t *__restrict__ out,
                                    int64_t *__restrict__ lens, int64_t *__restrict__ starts) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t start = mix64(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(r + 1))) % (ref_len - (uint64_t)len + 1);
    uint8_t *o = out ? out + offs[r] : nullptr;
    int64_t w = 0;
    for (int32_t p = 0; p < len; p++) {
        uint8_t c = ref ? ref[start + p] : synth_base(synth_seed, start + p);
        const uint64_t h = mix64((seed + 0x51ED270B1ull) ^ (0xD6E8FEB86659FD93ull * (uint64_t)(r * len + p + 1)));
        uint8_t extra = 0;
        bool drop = false;
        if ((uint32_t)h < thresh) {
            const uint32_t kind = (uint32_t)(h >> 32) & 0xFFFFu, pick = (uint32_t)(h >> 48);
            if (kind < ins_cut) extra = (uint8_t)"ACGT"[pick & 3];
            else if (kind < del_cut) drop = true;
            else { const int k = bwt_code(c); if (k >= 0) c = (uint8_t)"ACGT"[(k + 1 + (int)(pick % 3)) & 3]; }
        }
        for (int e = 0; e < 2; e++) {
            if (e == 0 ? drop : !extra) continue;
            uint8_t b = e == 0 ? c : extra;
            if (n_thresh) {
                const uint64_t hn = mix64((seed + 0x7F4A7C15ull) ^ (0xA24BAED4963EE407ull * (uint64_t)(2 * (r * len + p) + e + 1)));
                if ((uint32_t)hn < n_thresh) b = 'N';
            }
            if (o) o[w] = b;
            w++;
        }
    }
    if (!out) lens[r] = w;
    if (starts && !out) starts[r] = (int64_t)start;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; }
    int alloc(size_t bytes) {
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            bwt_set_error("builder: device allocation of %zu bytes failed", bytes);
            return BWT_ERR_NOMEM;
        }
        return BWT_OK;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

unsigned blocks_for(uint64_t n, unsigned t = 256) { return (unsigned)div_up_u64(n, t); }

// group structure of a sorted key array of m slots -> grp (head position per slot), tied flags
int compute_groups(const uint64_t *keys, uint64_t m, uint32_t *head, uint32_t *grp, uint32_t *tied, uint32_t *blk) {
    uint64_t nb = div_up_u64(m, 1024);
    heads_kernel<<<blocks_for(m), 256>>>(keys, m, head);
    block_last_head_kernel<<<blocks_for(nb), 256>>>(head, m, blk);
    carry_heads_kernel<<<1, 1>>>(blk, nb);
    groups_kernel<<<blocks_for(nb), 256>>>(head, m, blk, grp, tied);
    CU_TRY(cudaGetLastError());
    return BWT_OK;
}

}  // namespace

int build_sa_device(const uint8_t *d_text, uint64_t len, uint32_t **d_sa_out, uint8_t **d_L_out, uint32_t pointers[4]) {
    const uint64_t n = len + 1;
    if (len < 4 || n >= 0xFFFFFFFFull - (1u << 20)) { bwt_set_error("builder: %llu bases outside the supported range", (unsigned long long)len); return BWT_ERR_LIMIT; }
    // symbol counts -> pointers (first 1-based row whose F is C / G / T; fmi.erl:66-84)
    DevBuf cntb;
    BWT_TRY(cntb.alloc(4 * sizeof(unsigned long long)));
    CU_TRY(cudaMemset(cntb.p, 0, 4 * sizeof(unsigned long long)));
    count_kernel<<<148 * 8, 256>>>(d_text, len, cntb.as<unsigned long long>());
    unsigned long long cnt[4];
    CU_TRY(cudaMemcpy(cnt, cntb.p, sizeof cnt, cudaMemcpyDeviceToHost));
    if (cnt[0] + cnt[1] + cnt[2] + cnt[3] != len) { bwt_set_error("builder: text must be ACGT only"); return BWT_ERR_FORMAT; }
    if (!cnt[0] || !cnt[1] || !cnt[2] || !cnt[3]) {
        bwt_set_error("builder: every one of A,C,G,T must occur (fmi:get_index_pointers/1 walks off the index otherwise)");
        return BWT_ERR_FORMAT;
    }
    pointers[0] = (uint32_t)(2 + cnt[0]);
    pointers[1] = (uint32_t)(pointers[0] + cnt[1]);
    pointers[2] = (uint32_t)(pointers[1] + cnt[2]);
    pointers[3] = (uint32_t)n;

    // head / tied / grp alias the sort inputs (keys_a, idx_a), which are dead once a sort has produced keys_b / idx_b
    DevBuf keys_a, keys_b, idx_a, idx_b, rank, blk, tmp, scan_tmp, posa, posb;
    uint32_t *d_sa = nullptr;
    if (cudaMalloc(&d_sa, n * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); bwt_set_error("builder: SA allocation failed"); return BWT_ERR_NOMEM; }
    int st = BWT_OK;
    do {
        if ((st = keys_a.alloc(n * 8)) || (st = keys_b.alloc(n * 8)) || (st = idx_a.alloc(n * 4)) || (st = idx_b.alloc(n * 4)) ||
            (st = rank.alloc(n * 4)) ||
            (st = blk.alloc((div_up_u64(n, 1024) + 1) * 4)) || (st = scan_tmp.alloc(scan_tmp_words(n) * 4)))
            break;
        uint32_t *const a_head = keys_a.as<uint32_t>(), *const a_tied = keys_a.as<uint32_t>() + n, *const a_grp = idx_a.as<uint32_t>();
        key0_kernel<<<blocks_for(n), 256>>>(d_text, len, n, keys_a.as<uint64_t>(), idx_a.as<uint32_t>());
        // DoubleBuffer form: O(1) temporary storage (the pointer form would take another n*12 bytes); the 63-bit
        // key needs 8 passes, an even number, so the sorted data ends in the "b" halves -- checked, not assumed.
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<uint64_t> kbuf(keys_a.as<uint64_t>(), keys_b.as<uint64_t>());
        cub::DoubleBuffer<uint32_t> vbuf(idx_a.as<uint32_t>(), idx_b.as<uint32_t>());
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kbuf, vbuf, (uint64_t)n, 0, 64);
        tmp_bytes += 1024;
        if ((st = tmp.alloc(tmp_bytes))) break;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kbuf, vbuf, (uint64_t)n, 0, 63);
        if (e != cudaSuccess) { bwt_set_error("builder: radix sort failed: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; }
        if (kbuf.Current() != keys_b.as<uint64_t>()) {      // normalise: sorted keys in keys_b / idx_b
            CU_TRY(cudaMemcpy(keys_b.p, kbuf.Current(), n * 8, cudaMemcpyDeviceToDevice));
            CU_TRY(cudaMemcpy(idx_b.p, vbuf.Current(), n * 4, cudaMemcpyDeviceToDevice));
        }
        if ((st = compute_groups(keys_b.as<uint64_t>(), n, a_head, a_grp, a_tied, blk.as<uint32_t>()))) break;
        commit0_kernel<<<blocks_for(n), 256>>>(idx_b.as<uint32_t>(), a_grp, n, d_sa, rank.as<uint32_t>());
        // tied slots of round 0
        uint32_t m32 = 0;
        DevBuf totb;
        if ((st = totb.alloc(4))) break;
        // `tied` doubles as the scan input; keep the flags in `head` (no longer needed)
        cudaMemcpy(a_head, a_tied, n * 4, cudaMemcpyDeviceToDevice);
        if ((st = exclusive_scan_u32(a_tied, n, scan_tmp.as<uint32_t>(), totb.as<uint32_t>(), 0, nullptr))) break;
        cudaMemcpy(&m32, totb.p, 4, cudaMemcpyDeviceToHost);
        uint64_t m = m32;
        if (m) {
            if ((st = posa.alloc(m * 4)) || (st = posb.alloc(m * 4))) break;
            scatter_tied_kernel<<<blocks_for(n), 256>>>(a_head, a_tied, n, nullptr, posa.as<uint32_t>());
        }
        uint64_t h = 21;
        uint32_t *pos = posa.as<uint32_t>(), *pos_next = posb.as<uint32_t>();
        while (m > 0) {
            keyh_kernel<<<blocks_for(m), 256>>>(pos, m, d_sa, rank.as<uint32_t>(), n, h, keys_a.as<uint64_t>(), idx_a.as<uint32_t>());
            size_t need = 0;
            cub::DoubleBuffer<uint64_t> kb(keys_a.as<uint64_t>(), keys_b.as<uint64_t>());
            cub::DoubleBuffer<uint32_t> vb(idx_a.as<uint32_t>(), idx_b.as<uint32_t>());
            cub::DeviceRadixSort::SortPairs(nullptr, need, kb, vb, (uint64_t)m, 0, 64);
            if (need > tmp_bytes) { bwt_set_error("builder: radix temp grew"); st = BWT_ERR_NOMEM; break; }
            e = cub::DeviceRadixSort::SortPairs(tmp.p, need, kb, vb, (uint64_t)m, 0, 64);
            if (e != cudaSuccess) { bwt_set_error("builder: radix sort failed: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; }
            if (kb.Current() != keys_b.as<uint64_t>()) {
                CU_TRY(cudaMemcpy(keys_b.p, kb.Current(), m * 8, cudaMemcpyDeviceToDevice));
                CU_TRY(cudaMemcpy(idx_b.p, vb.Current(), m * 4, cudaMemcpyDeviceToDevice));
            }
            if ((st = compute_groups(keys_b.as<uint64_t>(), m, a_head, a_grp, a_tied, blk.as<uint32_t>()))) break;
            commith_kernel<<<blocks_for(m), 256>>>(pos, idx_b.as<uint32_t>(), a_grp, m, d_sa, rank.as<uint32_t>());
            cudaMemcpy(a_head, a_tied, m * 4, cudaMemcpyDeviceToDevice);
            if ((st = exclusive_scan_u32(a_tied, m, scan_tmp.as<uint32_t>(), totb.as<uint32_t>(), 0, nullptr))) break;
            cudaMemcpy(&m32, totb.p, 4, cudaMemcpyDeviceToHost);
            if (m32) scatter_tied_kernel<<<blocks_for(m), 256>>>(a_head, a_tied, m, pos, pos_next);
            std::swap(pos, pos_next);
            m = m32;
            h *= 2;
            if (h > 4 * n + 64) { bwt_set_error("builder: prefix doubling did not converge"); st = BWT_ERR_CUDA; break; }
        }
        if (st != BWT_OK) break;
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { bwt_set_error("builder: %s", cudaGetErrorString(e)); st = BWT_ERR_CUDA; break; }
    } while (0);
    if (st != BWT_OK) { cudaFree(d_sa); return st; }
    // free the big temporaries before allocating L
    keys_a.release();
    keys_b.release();
    uint8_t *d_L = nullptr;
    if (cudaMalloc(&d_L, n) != cudaSuccess) { cudaGetLastError(); cudaFree(d_sa); bwt_set_error("builder: L allocation failed"); return BWT_ERR_NOMEM; }
    bwt_kernel<<<blocks_for(n), 256>>>(d_text, d_sa, n, d_L);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d_sa); cudaFree(d_L); bwt_set_error("builder: %s", cudaGetErrorString(e)); return BWT_ERR_CUDA; }
    *d_sa_out = d_sa;
    *d_L_out = d_L;
    return BWT_OK;
}

// d_chunk: the .ref bytes on the device (kept as the window source); the text is the ACGT mapping of d_text_src
// (null: of d_chunk itself -- they differ when 'N's were substituted on the host)
static int build_from_device_chunk(uint8_t *d_chunk, const uint8_t *d_text_src, uint64_t len, int64_t shift, int device, bool already_acgt,
                                   bwt_index **out) {
    uint8_t *d_text = d_chunk;
    DevBuf mapped, flags;
    if (!already_acgt) {
        BWT_TRY(mapped.alloc(len));
        BWT_TRY(flags.alloc(8));
        CU_TRY(cudaMemset(flags.p, 0, 8));
        map_acgt_kernel<<<blocks_for(len), 256>>>(d_text_src ? d_text_src : d_chunk, len, mapped.as<uint8_t>(), flags.as<uint32_t>());
        uint32_t f[2];
        CU_TRY(cudaMemcpy(f, flags.p, 8, cudaMemcpyDeviceToHost));
        if (f[1]) { bwt_set_error("builder: %u symbols outside ACGT/IUPAC (fm_index.erl:136 raises)", f[1]); return BWT_ERR_FORMAT; }
        if (f[0]) {
            bwt_set_error("builder: %u 'N' in the chunk; the reference substitutes random:uniform(4) (fm_index.erl:131): pass an AS183 "
                          "state to bwt_build_n to reproduce it", f[0]);
            return BWT_ERR_FORMAT;
        }
        d_text = mapped.as<uint8_t>();
    }
    uint32_t *d_sa = nullptr;
    uint8_t *d_L = nullptr;
    uint32_t pointers[4];
    BWT_TRY(build_sa_device(d_text, len, &d_sa, &d_L, pointers));
    uint8_t *d_ref = nullptr;
    int st = index_upload_ref_device(d_chunk, len, &d_ref);
    if (st != BWT_OK) { cudaFree(d_sa); cudaFree(d_L); return st; }
    bwt_index *h = new bwt_index();
    h->device = device;
    st = index_finish_from_device(h, d_L, len + 1, d_sa, d_ref, len, pointers, shift);
    cudaFree(d_L);
    if (st != BWT_OK) { cudaFree(d_sa); cudaFree(d_ref); delete h; return st; }
    *out = h;
    return BWT_OK;
}

// random:uniform(4) of OTP's `random` module (AS183, Wichmann-Hill): the generator fm_index.erl:131 draws an 'N''s
// replacement from.  state = {A1,A2,A3}; the default seed of a fresh process is {3172, 9814, 20125}.  (From the OTP
// stdlib, which is not in the reference tree: parity of this substitution is unpinned there; the first draw of the
// default seed, 0.4435846174457203, is the value every Erlang shell prints.)
static int as183_uniform4(uint32_t st[3]) {
    st[0] = (st[0] * 171u) % 30269u;
    st[1] = (st[1] * 172u) % 30307u;
    st[2] = (st[2] * 170u) % 30323u;
    double r = (double)st[0] / 30269.0 + (double)st[1] / 30307.0 + (double)st[2] / 30323.0;
    r -= (double)(long long)r;
    return (int)(r * 4.0) + 1;
}

extern "C" int bwt_build_n(const uint8_t *chunk, uint64_t len, int64_t shift, int device, uint32_t as183_state[3], bwt_index **out) {
    if (!chunk || !out) { bwt_set_error("bwt_build: null argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    DevBuf d, dt;
    BWT_TRY(d.alloc(len));
    CU_TRY(cudaMemcpy(d.p, chunk, len, cudaMemcpyHostToDevice));
    if (as183_state && memchr(chunk, 'N', len)) {
        // element(random:uniform(4), {$A,$C,$G,$T}) per 'N', in text order (lists:map/2 walks the list from its head)
        std::vector<uint8_t> text(chunk, chunk + len);
        for (uint64_t k = 0; k < len; k++)
            if (text[k] == 'N') text[k] = (uint8_t)"ACGT"[as183_uniform4(as183_state) - 1];
        BWT_TRY(dt.alloc(len));
        CU_TRY(cudaMemcpy(dt.p, text.data(), len, cudaMemcpyHostToDevice));
        return build_from_device_chunk(d.as<uint8_t>(), dt.as<uint8_t>(), len, shift, device, false, out);
    }
    return build_from_device_chunk(d.as<uint8_t>(), nullptr, len, shift, device, false, out);
}

extern "C" int bwt_build(const uint8_t *chunk, uint64_t len, int64_t shift, int device, bwt_index **out) {
    return bwt_build_n(chunk, len, shift, device, nullptr, out);
}

extern "C" int bwt_build_synthetic(uint64_t len, uint64_t seed, int64_t shift, int device, bwt_index **out) {
    if (!out) { bwt_set_error("bwt_build_synthetic: null argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    DevBuf d;
    BWT_TRY(d.alloc(len));
    synth_kernel<<<blocks_for(len), 256>>>(d.as<uint8_t>(), 0, len, seed);
    CU_TRY(cudaGetLastError());
    return build_from_device_chunk(d.as<uint8_t>(), nullptr, len, shift, device, true, out);
}

extern "C" int bwt_build_synthetic_part(uint64_t start, uint64_t len, uint64_t seed, int64_t shift, int device, bwt_index **out) {
    if (!out) { bwt_set_error("bwt_build_synthetic_part: null argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    DevBuf d;
    BWT_TRY(d.alloc(len));
    synth_kernel<<<blocks_for(len), 256>>>(d.as<uint8_t>(), start, len, seed);
    CU_TRY(cudaGetLastError());
    return build_from_device_chunk(d.as<uint8_t>(), nullptr, len, shift, device, true, out);
}

extern "C" int bwt_sample_reads_edits(const bwt_index *h, const bwt_sample_opts *o, uint8_t *reads_out, int64_t reads_cap,
                                      int64_t *offsets_out, int64_t *starts_out, int64_t *total_bytes) {
    if (!o || !offsets_out || !total_bytes || o->n < 0 || o->read_len <= 0 || o->edit_rate < 0 || o->edit_rate > 1 ||
        o->ins_share < 0 || o->del_share < 0 || o->ins_share + o->del_share > 1 || o->n_rate < 0 || o->n_rate > 1) {
        bwt_set_error("bwt_sample_reads_edits: bad argument");
        return BWT_ERR_ARG;
    }
    const uint64_t ref_len = h ? h->ref_len : o->synth_len;
    if ((uint64_t)o->read_len > ref_len) { bwt_set_error("bwt_sample_reads_edits: reads longer than the reference"); return BWT_ERR_ARG; }
    const int64_t n = o->n;
    *total_bytes = 0;
    offsets_out[0] = 0;
    if (n == 0) return BWT_OK;
    BWT_TRY(index_select_device(h ? h->device : o->device));
    auto frac32 = [](double f) { double t = f * 4294967296.0; return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t; };
    const uint32_t thresh = frac32(o->edit_rate), n_thresh = frac32(o->n_rate);
    const uint32_t ins_cut = (uint32_t)(o->ins_share * 65536.0), del_cut = (uint32_t)((o->ins_share + o->del_share) * 65536.0);
    const uint8_t *ref = h ? (const uint8_t *)h->d_ref + BWT_REF_EXT : nullptr;
    DevBuf lens, offs, st, tmp, out;
    BWT_TRY(lens.alloc((size_t)(n + 1) * sizeof(int64_t)));
    BWT_TRY(offs.alloc((size_t)(n + 1) * sizeof(int64_t)));
    if (starts_out) BWT_TRY(st.alloc((size_t)n * sizeof(int64_t)));
    CU_TRY(cudaMemset(lens.p, 0, (size_t)(n + 1) * sizeof(int64_t)));
    sample_edits_kernel<<<blocks_for((uint64_t)n, 128), 128>>>(ref, o->synth_seed, ref_len, n, o->read_len, o->seed, thresh, ins_cut, del_cut,
                                                             n_thresh, nullptr, nullptr, lens.as<int64_t>(),
                                                             starts_out ? st.as<int64_t>() : nullptr);
    CU_TRY(cudaGetLastError());
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens.as<int64_t>(), offs.as<int64_t>(), (int64_t)(n + 1));
    BWT_TRY(tmp.alloc(tmp_bytes));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, lens.as<int64_t>(), offs.as<int64_t>(), (int64_t)(n + 1));
    if (e != cudaSuccess) { bwt_set_error("bwt_sample_reads_edits: scan failed: %s", cudaGetErrorString(e)); return BWT_ERR_CUDA; }
    CU_TRY(cudaMemcpy(offsets_out, offs.p, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    const int64_t total = offsets_out[n];
    *total_bytes = total;
    if (starts_out) CU_TRY(cudaMemcpy(starts_out, st.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (!reads_out || total > reads_cap) { bwt_set_error("bwt_sample_reads_edits: reads need %lld bytes", (long long)total); return BWT_ERR_NOSPACE; }
    lens.release(); tmp.release();
    BWT_TRY(out.alloc((size_t)total + 16));
    sample_edits_kernel<<<blocks_for((uint64_t)n, 128), 128>>>(ref, o->synth_seed, ref_len, n, o->read_len, o->seed, thresh, ins_cut, del_cut,
                                                             n_thresh, offs.as<int64_t>(), out.as<uint8_t>(), nullptr, nullptr);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(reads_out, out.p, (size_t)total, cudaMemcpyDeviceToHost));
    return BWT_OK;
}

extern "C" int bwt_sample_reads(const bwt_index *h, int64_t n, int32_t read_len, uint64_t seed, double sub_rate,
                                uint8_t *reads_out, int64_t *starts_out) {
    if (!h || !reads_out || n < 0 || read_len <= 0 || sub_rate < 0 || sub_rate > 1 || (uint64_t)read_len > h->ref_len) {
        bwt_set_error("bwt_sample_reads: bad argument");
        return BWT_ERR_ARG;
    }
    if (n == 0) return BWT_OK;
    BWT_TRY(index_select_device(h->device));
    DevBuf out, st;
    BWT_TRY(out.alloc((size_t)n * read_len));
    if (starts_out) BWT_TRY(st.alloc((size_t)n * sizeof(int64_t)));
    double t = sub_rate * 4294967296.0;
    uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    sample_reads_kernel<<<blocks_for((uint64_t)n * read_len), 256>>>((const uint8_t *)h->d_ref + BWT_REF_EXT, h->ref_len, n, read_len,
                                                                    seed, thresh, out.as<uint8_t>(),
                                                                    starts_out ? st.as<int64_t>() : nullptr);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(reads_out, out.p, (size_t)n * read_len, cudaMemcpyDeviceToHost));
    if (starts_out) CU_TRY(cudaMemcpy(starts_out, st.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return BWT_OK;
}
