"""Mirror of seeds.erl: generate_fs/3 (seeds.erl:11-47) and the `fs` module it writes, on the GPU index.

The reference prints an Erlang source file (fs.erl) holding one `matches/Len` clause per pattern of gen_pattern/5 and
a window scanner `find_seeds/1`; a caller would compile it and run `fs:find_seeds(Ref_Seq)`.  Here generate_fs/3
returns the "module" as an object whose find_seeds/1 takes the FM index of the chunk instead of its text; the
patterns are never materialised (inexact.cu walks gen_pattern/5's recursion on the index).
"""
from . import api
from .errors import ReferenceCrash


class Fs:
    """The module seeds:generate_fs(Qsec, Len, Defects) writes."""

    def __init__(self, qsec: bytes, length: int, defects: int):
        self.qsec, self.length, self.defects = bytes(qsec), int(length), int(defects)

    def find_seeds(self, fm, max_raw_hits: int = 1024):
        """fs:find_seeds(Ref_Seq) -> [N] (1-based window starts, descending), Ref_Seq = the text of `fm`."""
        return find_seeds_batch(fm, [self.qsec], self.length, self.defects, max_raw_hits)[0]


def generate_fs(qsec: bytes, length: int, defects: int) -> Fs:
    if length < 2:
        raise ReferenceCrash("badmatch at seeds.erl:35: Len must be at least 2")
    return Fs(qsec, length, defects)


def find_seeds_batch(fm, qsecs, length: int, defects: int, max_raw_hits: int = 1024):
    """fs:find_seeds/1 of generate_fs(Qsec, Len, Defects) for every Qsec -> list of lists"""
    st, so, seeds, _ = fm.fs_find_seeds(qsecs, None, length, defects, max_raw_hits)
    out = []
    for r in range(len(qsecs)):
        if st[r] == api.FS_SHORT:
            raise ReferenceCrash("fs.erl of query %d has no usable clause or does not compile: Qsec shorter than "
                                 "Len + Defects (seeds.erl:51-55)" % r)
        if st[r] == api.FS_LIMIT:
            raise ValueError("query %d: more matching windows than max_raw_hits=%d" % (r, max_raw_hits))
        out.append(seeds[so[r]:so[r + 1]].tolist())
    return out
