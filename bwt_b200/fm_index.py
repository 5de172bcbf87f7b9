"""Mirror of fm_index.erl's online half: get_index/2 (fm_index.erl:154-156) -> a GPU-resident index.

``get_index(Chrom, J, fm_indices=...)`` returns ``(Meta, FM)`` like the reference: Meta is the proplist
``[("pointers", (Pc,Pg,Pt,Last)), ("shift", Shift)]`` and FM the opaque device index (``api.Index``) that takes the
place of the row tuple.  ``make_index`` mirrors fm_index:make_index/4 (device builder + reference-format writer).
"""
import os

from . import api


def get_index(chrom: str, part: int, fm_indices: str = "fm_indices", device: int = 0):
    base = os.path.join(fm_indices, "%s_p%d" % (chrom, part))
    fm = api.Index.open(base + ".fm", base + ".ref", device)
    i = fm.info()
    return [("pointers", tuple(i.pointers)), ("shift", int(i.shift))], fm


def make_index(chunk: bytes, chrom: str, part: int, shift: int, fm_indices: str = "fm_indices", device: int = 0):
    """fm_index.erl:127-141: write <Chrom>_p<J>.ref (raw chunk) and <Chrom>_p<J>.fm (term_to_binary({Meta,FM}))."""
    base = os.path.join(fm_indices, "%s_p%d" % (chrom, part))
    with api.Index.build(chunk, shift=shift, device=device) as fm:
        fm.save(base + ".fm", base + ".ref")
    return base + ".fm", base + ".ref"


def chunk_plan(body_len: int, nbr=None, overlap: int = 2):
    """fm_index.erl:113-124: [(Start (1-based), length)] for J = 1..Nbr; Nbr = 1 + ((Len+2) bsr 24) by default."""
    if nbr is None:
        nbr = 1 + ((body_len + overlap) >> 24)
    size = (body_len + nbr * overlap) // nbr
    return [(1 + (j - 1) * (size - overlap), max(0, min(size, body_len - (1 + (j - 1) * (size - overlap)) + 1)))
            for j in range(1, nbr + 1)]
