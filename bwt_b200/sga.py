"""Mirror of sga.erl: sga/6 (sga.erl:17-64) on the GPU index."""
from . import api
from .errors import LibraryLimit, ReferenceCrash


def sga(fm, pc, pg, pt, last, qseq: bytes):
    """-> [(S, D)] descending by position, [] when fewer than two subsequences produced seeds."""
    return sga_batch(fm, [qseq])[0]


def sga_batch(fm, qseqs):
    st, clo, S, D = fm.sga(qseqs)
    out = []
    for r in range(len(qseqs)):
        if st[r] == api.READ_CRASH:
            raise ReferenceCrash("bwa:find_seeds crashes on a subsequence of read %d" % r)
        if st[r] == api.READ_LIMIT:
            raise LibraryLimit("read %d exceeds BWT_MAX_READ_LEN" % r)
        out.append(list(zip(S[clo[r]:clo[r + 1]].tolist(), D[clo[r]:clo[r + 1]].tolist())))
    return out
