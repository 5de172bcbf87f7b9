// Seeding kernels: batched backward search, and the fused per-read
//   skip_Ns -> find_seeds -> locate -> sort -> vote        (sga:sga/6, sga.erl:17-76)
// with one cooperative group of G lanes per read.  Seed ends never leave the SM: they are located into
// shared memory, sorted there (bitonic) and voted into {S,D} clusters; only clusters go to HBM.
#pragma once
#include "common.cuh"

struct Cluster {        // one element of sga's result list, 16 B
    uint32_t read;      // read index inside the batch
    uint32_t s;         // S: 0-based end (exclusive) of the read in the chunk
    int32_t d;          // D: sum of gaps of the voting chain
    uint32_t rank;      // position in the reference's (descending) result list
};

// Clusters are appended to SGA_SUBLISTS sub-lists (one region of the cluster array and one cursor on its own 128-byte
// line each, chosen by block): ten million appends to ONE cursor serialise in a single L2 slice (1.75 ns each: the
// whole kernel ran at the speed of that atomic).  launch_sga_finish packs the regions into the dense list.
constexpr int SGA_SUBLISTS = 64;

struct SeedCounters {   // device-side counters of one run
    unsigned long long searches, lf_steps, hbm_lf_steps, located, clusters;
    uint32_t n_clusters;      // clusters appended (sum over the sub-lists; set by launch_sga_finish)
    uint32_t n_slow;          // reads whose seed ends did not fit the shared-memory budget
    uint32_t cluster_overflow;
    uint32_t scratch_overflow;   // a deferred read did not fit the slow kernel's scratch
    unsigned long long scratch_used;
    uint32_t n_pairs;            // what the kernels behind the seeding stage work on: n_clusters, or 0 when a capacity overflowed
    uint32_t abort;              // a capacity overflowed: the rest of the run is skipped on the device, the host grows and reruns
    uint32_t pad2[2];
    uint32_t dense_base[SGA_SUBLISTS];          // first dense slot of each sub-list (launch_sga_finish)
    uint32_t sub_n[SGA_SUBLISTS * 32];          // cursor of sub-list r at [32 r]
};

struct SlowRead { uint32_t read; uint32_t seeds; };

constexpr int SGA_MAX_TASKS = 128;    // subsequences per read => BWT_MAX_READ_LEN = 8*127 + 15

// launchers (seed.cu)
int launch_search(const DevIndex &ix, const uint8_t *d_subs, const int64_t *d_offs, int64_t n, int32_t fixed_len,
                  int32_t *d_status, uint32_t *d_sp, uint32_t *d_ep, int32_t *d_n, SeedCounters *d_ctr,
                  cudaStream_t stream);
int launch_sga(const DevIndex &ix, const DevReads &R, int64_t n_reads, int max_read_len,
               int32_t *d_read_status, uint32_t *d_ncl, uint32_t *d_cl_base, Cluster *d_clusters, uint32_t cl_cap,
               SlowRead *d_slow, uint32_t slow_cap, SeedCounters *d_ctr, cudaStream_t stream);
// Packs the sub-lists of d_sparse into d_dense, turns cl_base into dense indices and sets d_ctr->n_clusters.
int launch_sga_finish(const Cluster *d_sparse, Cluster *d_dense, uint32_t cl_cap, uint32_t *d_cl_base, const uint32_t *d_ncl,
                      int64_t n_reads, uint32_t slow_cap, SeedCounters *d_ctr, cudaStream_t stream, uint32_t *launches);
int launch_sga_slow(const DevIndex &ix, const DevReads &R, const SlowRead *d_slow, uint32_t slow_cap,
                    uint32_t *d_scratch, uint64_t scratch_cap, int32_t *d_read_status,
                    uint32_t *d_ncl, uint32_t *d_cl_base, Cluster *d_clusters, uint32_t cl_cap, SeedCounters *d_ctr,
                    cudaStream_t stream);
