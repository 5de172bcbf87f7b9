"""Mirror of fastq.erl (fastq.erl:5-40) over the native parser ``bwt_fastq_parse``: same record rule, same
reverse-order batches, but the bases land directly in the packed layout the aligner takes."""
import ctypes as C

import numpy as np

from ._lib import check, lib


def parse(text, max_reads=None, final=True):
    """-> (reads uint8, offsets int64[n+1], names [bytes], consumed bytes); records in FILE order."""
    buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    if max_reads is None:
        max_reads = int(np.count_nonzero(buf == 10)) // 2 + 2
    reads = np.empty(max(len(buf), 1), np.uint8)
    offs = np.zeros(max_reads + 1, np.int64)
    ns = np.zeros(max(max_reads, 1), np.int64)
    nl = np.zeros(max(max_reads, 1), np.int32)
    n, used = C.c_int64(), C.c_int64()
    check(lib().bwt_fastq_parse(buf.ctypes.data if len(buf) else None or reads.ctypes.data, len(buf), int(final), max_reads,
                                reads.ctypes.data, len(reads), offs.ctypes.data, ns.ctypes.data, nl.ctypes.data,
                                C.byref(n), C.byref(used)))
    n = int(n.value)
    raw = buf.tobytes()
    names = [raw[ns[k]:ns[k] + nl[k]] for k in range(n)]
    return reads[:offs[n]].copy(), offs[:n + 1].copy(), names, int(used.value)


class Reader:
    """fastq:read_seqs/2 on an in-memory FASTQ: batches of N records in REVERSE file order (fastq.erl:20-40)."""

    def __init__(self, text: bytes):
        self.reads, self.offs, self.names, _ = parse(text)
        self.pos = 0

    def read_seqs(self, n: int):
        """-> ("ok", [(name, seq)]) | ("eof", [...]) | "eof" """
        total = len(self.names)
        if self.pos >= total:
            return "eof"
        hi = min(total, self.pos + n)
        batch = [(self.names[k], self.reads[self.offs[k]:self.offs[k + 1]].tobytes()) for k in range(hi - 1, self.pos - 1, -1)]
        tag = "ok" if hi - self.pos == n else "eof"
        self.pos = hi
        return tag, batch
