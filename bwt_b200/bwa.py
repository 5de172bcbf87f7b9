"""Mirror of bwa.erl: find_seeds/6 (bwa.erl:12-45) on the GPU index."""
from . import api
from .errors import ReferenceCrash

NO_SEEDS = "no_seeds"
TOO_MANY_SEEDS = "too_many_seeds"


def find_seeds(fm, pc, pg, pt, last, subseq: bytes):
    """-> "no_seeds" | "too_many_seeds" | [SA(Ep)+N, ..., SA(Sp)+N]   (pointer arguments kept for signature parity)"""
    return find_seeds_batch(fm, [subseq])[0]


def find_seeds_batch(fm, subseqs):
    st, sp, ep, n = fm.search(subseqs)
    out = []
    for k in range(len(subseqs)):
        if st[k] == api.SEARCH_CRASH:
            raise ReferenceCrash("case_clause / badarg in bwa:find_seeds for %r" % (subseqs[k],))
        if st[k] == api.NO_SEEDS:
            out.append(NO_SEEDS)
        elif st[k] == api.TOO_MANY_SEEDS:
            out.append(TOO_MANY_SEEDS)
        else:
            sa = fm.locate(int(sp[k]), int(ep[k]) - int(sp[k]) + 1)
            out.append([int(x) + int(n[k]) for x in sa[::-1]])      # fmi.erl:89-100
    return out
