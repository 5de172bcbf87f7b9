"""Mirror of sw.erl: sw/2 (sw.erl:47-160) on the GPU."""
from . import api

NO_MATCH = "no_match"


def sw(qseq: bytes, ref: bytes, device: int = 0):
    """-> "no_match" | (Score, CigarString)"""
    return sw_batch([(qseq, ref)], device)[0]


def sw_batch(pairs, device: int = 0):
    q, qo = api.pack([p[0] for p in pairs])
    r, ro = api.pack([p[1] for p in pairs])
    st, sc, cig = api.sw_batch(q, qo, r, ro, device)
    return [(int(sc[k]), cig[k].decode()) if st[k] == api.HIT_MATCH else NO_MATCH for k in range(len(pairs))]
