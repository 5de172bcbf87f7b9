// Host-side parser/writer of the reference's .fm container (see etf.cpp).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

struct EtfIndex {
    std::vector<uint8_t> L;     // ASCII over "$ACGT"
    std::vector<uint32_t> SA;   // 0-based suffix starts
    uint32_t pointers[4] = {0, 0, 0, 0};
    int64_t shift = 0;
};

bool etf_parse_fm(const uint8_t *data, size_t size, EtfIndex &out, std::string &err);
void etf_write_fm(const uint8_t *L, const uint32_t *SA, uint64_t n, const uint32_t pointers[4], int64_t shift,
                  std::vector<uint8_t> &out);
