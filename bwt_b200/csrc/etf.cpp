// Reader / writer for the reference's on-disk index:  <Chrom>_pJ.fm = term_to_binary({Meta, FM})
// (fm_index.erl:139-141, read back by fm_index:get_index/2, fm_index.erl:154-156).
//
//   Meta = [{pointers,{Pc,Pg,Pt,Last}},{shift,Shift}]      FM = {{F,L,I,SA}, ...}  (n-tuple of 4-tuples)
//
// Erlang external term format, uncompressed, as OTP's term_to_binary/1 emits it for this shape:
//   131, SMALL_TUPLE(104) 2, LIST(108) 2, ..., NIL(106), LARGE_TUPLE(105) n | SMALL_TUPLE(104) n, rows...
// Integers: SMALL_INTEGER(97) u8, INTEGER(98) i32 big-endian, SMALL_BIG(110) for |v| >= 2^31.
// Only L, SA, pointers and shift are kept: F and I are functions of them.
#include "etf.hpp"

#include <new>
#include <stdio.h>
#include <string.h>

namespace {

struct Rd {
    const uint8_t *p, *end;
    bool ok = true;
    uint8_t u8() {
        if (p >= end) { ok = false; return 0; }
        return *p++;
    }
    uint32_t be32() {
        if (end - p < 4) { ok = false; p = end; return 0; }
        uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        p += 4;
        return v;
    }
    uint16_t be16() {
        if (end - p < 2) { ok = false; p = end; return 0; }
        uint16_t v = (uint16_t)((p[0] << 8) | p[1]);
        p += 2;
        return v;
    }
    bool integer(int64_t &v) {
        uint8_t tag = u8();
        if (tag == 97) { v = u8(); return ok; }
        if (tag == 98) { v = (int32_t)be32(); return ok; }
        if (tag == 110) {
            uint8_t k = u8(), sign = u8();
            if (k > 8 || end - p < k) { ok = false; return false; }
            uint64_t m = 0;
            for (int i = 0; i < k; i++) m |= (uint64_t)p[i] << (8 * i);
            p += k;
            v = sign ? -(int64_t)m : (int64_t)m;
            return ok;
        }
        ok = false;
        return false;
    }
    // ATOM_EXT (100, what OTP 17 wrote: the shipped fixture), SMALL_ATOM_EXT (115), and the UTF-8 forms current OTP
    // (26+) writes: ATOM_UTF8_EXT (118), SMALL_ATOM_UTF8_EXT (119)
    bool atom(const char *want) {
        const uint8_t tag = u8();
        size_t k;
        if (tag == 100 || tag == 118) k = be16();
        else if (tag == 115 || tag == 119) k = u8();
        else { ok = false; return false; }
        size_t wl = strlen(want);
        if (!ok || (size_t)(end - p) < k || k != wl || memcmp(p, want, k) != 0) { ok = false; return false; }
        p += k;
        return true;
    }
    bool small_tuple(unsigned arity) { return u8() == 104 && u8() == arity && ok; }
};

void put_int(std::vector<uint8_t> &o, int64_t v) {
    if (v >= 0 && v <= 255) { o.push_back(97); o.push_back((uint8_t)v); return; }
    if (v >= INT32_MIN && v <= INT32_MAX) {
        uint32_t u = (uint32_t)(int32_t)v;
        o.push_back(98); o.push_back(u >> 24); o.push_back(u >> 16); o.push_back(u >> 8); o.push_back(u);
        return;
    }
    uint64_t m = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
    uint8_t d[8]; int k = 0;
    while (m) { d[k++] = (uint8_t)m; m >>= 8; }
    o.push_back(110); o.push_back((uint8_t)k); o.push_back(v < 0 ? 1 : 0);
    for (int i = 0; i < k; i++) o.push_back(d[i]);
}

}  // namespace

bool etf_parse_fm(const uint8_t *data, size_t size, EtfIndex &out, std::string &err) {
    Rd r{data, data + size};
    if (r.u8() != 131) { err = "not an Erlang external term (missing version byte 131)"; return false; }
    if (!r.small_tuple(2)) { err = "top term is not a 2-tuple {Meta,FM}"; return false; }
    if (r.u8() != 108 || r.be32() != 2) { err = "Meta is not a 2-element list"; return false; }
    int64_t v;
    if (!r.small_tuple(2) || !r.atom("pointers") || !r.small_tuple(4)) { err = "Meta lacks {pointers,{Pc,Pg,Pt,Last}}"; return false; }
    for (int k = 0; k < 4; k++) {
        if (!r.integer(v) || v < 0 || v > 0xFFFFFFFFll) { err = "bad pointer value"; return false; }
        out.pointers[k] = (uint32_t)v;
    }
    if (!r.small_tuple(2) || !r.atom("shift") || !r.integer(v)) { err = "Meta lacks {shift,S}"; return false; }
    out.shift = v;
    if (r.u8() != 106) { err = "Meta list is not nil-terminated"; return false; }
    uint64_t n;
    uint8_t tag = r.u8();
    if (tag == 105) n = r.be32();
    else if (tag == 104) n = r.u8();
    else { err = "FM is not a tuple"; return false; }
    if (!r.ok) { err = "truncated header"; return false; }
    // the smallest row is 10 bytes (104 4 + four SMALL_INTEGERs): a count beyond that is a truncated or corrupt file,
    // not a reason to ask for gigabytes
    if (n > (uint64_t)(r.end - r.p) / 10) { err = "row count exceeds the file size (truncated or corrupt .fm)"; return false; }
    try {
        out.L.resize(n);
        out.SA.resize(n);
    } catch (const std::bad_alloc &) {
        err = "out of host memory for the index rows";
        return false;
    }
    for (uint64_t e = 0; e < n; e++) {
        int64_t F, L, I, SA;
        if (!r.small_tuple(4) || !r.integer(F) || !r.integer(L) || !r.integer(I) || !r.integer(SA)) {
            char buf[96];
            snprintf(buf, sizeof buf, "bad FM row %llu", (unsigned long long)(e + 1));
            err = buf;
            return false;
        }
        if (L < 0 || L > 255 || SA < 0 || SA > 0xFFFFFFFFll) { err = "FM row field out of range"; return false; }
        out.L[e] = (uint8_t)L;
        out.SA[e] = (uint32_t)SA;
    }
    if (r.p != r.end) { err = "trailing bytes after the FM tuple"; return false; }
    return true;
}

void etf_write_fm(const uint8_t *L, const uint32_t *SA, uint64_t n, const uint32_t pointers[4], int64_t shift,
                  std::vector<uint8_t> &o) {
    o.clear();
    o.reserve(64 + n * 16);
    const uint8_t head[] = {131, 104, 2, 108, 0, 0, 0, 2, 104, 2, 100, 0, 8, 'p', 'o', 'i', 'n', 't', 'e', 'r', 's', 104, 4};
    o.insert(o.end(), head, head + sizeof head);
    for (int k = 0; k < 4; k++) put_int(o, pointers[k]);
    const uint8_t sh[] = {104, 2, 100, 0, 5, 's', 'h', 'i', 'f', 't'};
    o.insert(o.end(), sh, sh + sizeof sh);
    put_int(o, shift);
    o.push_back(106);
    if (n <= 255) { o.push_back(104); o.push_back((uint8_t)n); }
    else { o.push_back(105); o.push_back(n >> 24); o.push_back(n >> 16); o.push_back(n >> 8); o.push_back(n); }
    // F of row e (1-based): '$' for row 1, then A / C / G / T by the pointer ranges (fmi.erl:66-84);
    // I = 0-based row of the occ(c,e)-th row whose F is c (bwt.erl:105-120, fmi.erl:43-62); 0 for the '$' row.
    uint64_t next[4] = {1, (uint64_t)pointers[0] - 1, (uint64_t)pointers[1] - 1, (uint64_t)pointers[2] - 1};
    for (uint64_t k = 0; k < n; k++) {
        uint64_t e = k + 1;
        uint8_t F = e == 1 ? '$' : e < pointers[0] ? 'A' : e < pointers[1] ? 'C' : e < pointers[2] ? 'G' : 'T';
        int c = L[k] == 'A' ? 0 : L[k] == 'C' ? 1 : L[k] == 'G' ? 2 : L[k] == 'T' ? 3 : -1;
        int64_t I = c < 0 ? 0 : (int64_t)next[c]++;
        o.push_back(104); o.push_back(4);
        put_int(o, F); put_int(o, L[k]); put_int(o, I); put_int(o, SA[k]);
    }
}
