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


def _contigs(data: bytes):
    """find_chromosome/2 + read_chromosome/2 (fm_index.erl:71-91): yields (name, body_offset, body_len) per '>' record;
    name = first token of the header line split on blanks and newlines."""
    pos, n = 0, len(data)
    cur = None
    while pos < n:
        nl = data.find(b"\n", pos)
        end = n if nl < 0 else nl + 1
        if data[pos:pos + 1] == b">":
            if cur is not None:
                yield cur[0], cur[1], pos - cur[1]
            toks = data[pos + 1:end].replace(b"\n", b" ").split()
            cur = (toks[0].decode("latin1") if toks else "", end)
        pos = end
    if cur is not None:
        yield cur[0], cur[1], n - cur[1]


def contig_body(raw: bytes):
    """build_fm/4 (fm_index.erl:93-111): trim N / CR / LF at both ends (left-trimmed N's are the shift), then drop every
    CR and LF -> (Shift, Body)."""
    skip = (78, 10, 13)
    hi = len(raw)
    while hi > 0 and raw[hi - 1] in skip:
        hi -= 1
    lo = shift = 0
    while lo < hi and raw[lo] in skip:
        shift += raw[lo] == 78
        lo += 1
    return shift, raw[lo:hi].replace(b"\n", b"").replace(b"\r", b"")


def make_indices(genome_file: str, chromo_ls="all", fm_indices: str = "fm_indices", device: int = 0):
    """fm_index:make_indices/2 (fm_index.erl:28-64): index every listed contig of a FASTA, chunked at 2^24 bp with a
    2 bp overlap (fm_index.erl:113-124); -> [(chrom, J, shift, fm_path, ref_path)]."""
    with open(genome_file, "rb") as f:
        data = f.read()
    os.makedirs(fm_indices, exist_ok=True)
    made = []
    for name, off, ln in _contigs(data):
        if chromo_ls != "all" and name not in chromo_ls:
            continue
        shift, body = contig_body(data[off:off + ln])
        for j, (start, length) in enumerate(chunk_plan(len(body)), start=1):
            chunk = body[start - 1:start - 1 + length]
            fm_path, ref_path = make_index(chunk, name, j, shift + start, fm_indices, device)
            made.append((name, j, shift + start, fm_path, ref_path))
    return made
