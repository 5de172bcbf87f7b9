// FASTQ ingest for the aligner: fastq:read_seq/1 (fastq.erl:5-18) over a text buffer, straight
// into the packed layout the C ABI takes (concatenated bases + offsets).  Host code; no GPU work here.
//
// Record rule of the reference, kept exactly: a line that starts with '@' is a header -- the name is the line minus
// the '@' and minus its LAST character (lists:droplast/1: the newline, or a real character when the file ends without
// one); the next line is the sequence (again minus its last character); the two lines after it are skipped whatever
// they hold.  Lines that do not start with '@' are skipped one at a time.  A header that is the last line of the
// buffer is a badmatch in the reference ({ok,SeqData} = eof): reported as BWT_ERR_FORMAT when `final` is set,
// otherwise the record is left unconsumed for the next buffer.
#include <string.h>

#include "common.cuh"

namespace {
// [p, line_end) is the line without its '\n'; returns the start of the next line, or nullptr if the buffer ends first
inline const uint8_t *next_line(const uint8_t *p, const uint8_t *end, const uint8_t **line_end, bool *had_newline) {
    const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
    if (nl) { *line_end = nl; *had_newline = true; return nl + 1; }
    *line_end = end; *had_newline = false;
    return end;
}
}  // namespace

extern "C" int bwt_fastq_parse(const uint8_t *text, int64_t len, int final, int64_t max_reads, uint8_t *reads_out,
                               int64_t reads_cap, int64_t *offsets, int64_t *name_start, int32_t *name_len,
                               int64_t *n_reads, int64_t *consumed) {
    if (!text || len < 0 || max_reads < 0 || !offsets || !n_reads || !consumed || (!reads_out && reads_cap > 0)) {
        bwt_set_error("bwt_fastq_parse: bad argument");
        return BWT_ERR_ARG;
    }
    const uint8_t *p = text, *end = text + len;
    int64_t n = 0, used = 0;
    offsets[0] = 0;
    int status = BWT_OK;
    while (p < end && n < max_reads) {
        const uint8_t *rec = p, *le;
        bool nl;
        const uint8_t *q = next_line(p, end, &le, &nl);
        if (!nl && !final) break;                                   // partial line: wait for more text
        if (*p != '@') { p = q; continue; }                         // {ok, _} -> read_seq(Dev)
        // header: name = line minus '@' minus its last character
        int64_t hs = (p + 1) - text;
        int64_t hl = (nl ? (le + 1) : le) - (p + 1) - 1;            // droplast of "name\n" (or of "name" at EOF)
        if (hl < 0) hl = 0;
        if (q >= end) {                                             // no sequence line
            if (final) { bwt_set_error("FASTQ ends after a header line (badmatch in fastq:read_seq/1)"); status = BWT_ERR_FORMAT; }
            p = rec;
            break;
        }
        const uint8_t *s0 = q, *se;
        bool nl2;
        const uint8_t *q2 = next_line(q, end, &se, &nl2);
        if (!nl2 && !final) { p = rec; break; }
        int64_t sl = (nl2 ? (se + 1) : se) - s0 - 1;                // droplast
        if (sl < 0) sl = 0;
        // delimiter and quality lines: read and ignored (eof there is fine for the reference)
        const uint8_t *q3 = q2, *q4 = q2, *tmp;
        bool nl3 = true, nl4 = true;
        if (q2 < end) q3 = next_line(q2, end, &tmp, &nl3);
        q4 = q3;
        if (q3 < end) q4 = next_line(q3, end, &tmp, &nl4);
        if (!final && (q2 >= end || q3 >= end || !nl3 || !nl4)) { p = rec; break; }   // record not complete yet
        if (used + sl > reads_cap) { bwt_set_error("bwt_fastq_parse: read buffer too small"); status = BWT_ERR_NOSPACE; p = rec; break; }
        memcpy(reads_out + used, s0, (size_t)sl);
        used += sl;
        if (name_start) name_start[n] = hs;
        if (name_len) name_len[n] = (int32_t)hl;
        offsets[++n] = used;
        p = q4;
    }
    *n_reads = n;
    *consumed = p - text;
    return status;
}
