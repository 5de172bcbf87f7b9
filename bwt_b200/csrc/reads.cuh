// Device-side read store helpers (reads.cu): geometry (lengths, word offsets) and 2-bit packing of a batch.
#pragma once
#include "common.cuh"

// lengths + word offsets from ASCII offsets already on the device (rebased to 0); d_info[0] = atomicMax of the lengths
// is NOT computed here (the host knows the offsets); d_info[1] is set when the offsets decrease
int reads_geom_from_ascii(const int64_t *d_aoff, int64_t n, uint32_t *d_len, uint32_t *d_woff, uint32_t *d_scan_tmp,
                          uint32_t *d_info, cudaStream_t stream, uint32_t *launches);
int reads_geom_fixed(int64_t n, uint32_t fixed_len, uint32_t *d_len, uint32_t *d_woff, cudaStream_t stream, uint32_t *launches);
int reads_geom_from_len16(const uint16_t *d_len16, int64_t n, uint32_t *d_len, uint32_t *d_woff, uint32_t *d_scan_tmp,
                          cudaStream_t stream, uint32_t *launches);
// ASCII -> packed words + exception mask (one bit per read)
int reads_pack_device(const uint8_t *d_ascii, const int64_t *d_aoff, int64_t aoff_base, int64_t n, const uint32_t *d_len, const uint32_t *d_woff,
                      uint32_t *d_pk, uint32_t *d_exc, cudaStream_t stream, uint32_t *launches);
// exception mask from the list of flagged reads (host-packed batches)
int reads_exc_scatter(const uint32_t *d_exc_idx, uint32_t n_exc, uint32_t base, int64_t n, uint32_t *d_exc, cudaStream_t stream, uint32_t *launches);
