"""Mirror of worker_bwt.erl's inner loop (worker_bwt.erl:73-108): one list of {Qname,Qseq} in, the messages the
worker would cast to the master out, in the reference's emission order."""
from . import api
from .errors import LibraryLimit, ReferenceCrash

REF_EXTENSION_LEN = 200   # worker_bwt.erl:17


def otp_sort_top(cigars):
    """hd(lists:sort(fun({{R1,_},_,_},{{R2,_},_,_}) -> R1 > R2 end, Cigars)) (worker_bwt.erl:104).
    The comparator is strict, so among equal scores OTP's merge sort decides; as recalled from the OTP source
    (not in the reference tree -> parity UNPINNED) a run of equal elements comes out reversed, i.e.
    the LAST of the best-scoring candidates wins.  Kept on the host, exactly where the reference keeps it."""
    best = max(c[0][0] for c in cigars)
    return [c for c in cigars if c[0][0] == best][-1]


def process_list(fm, padded_ref: bytes, qseq_list):
    """qseq_list: [(Qname, Qseq bytes)] -> list of messages: "no_cigar" | ("cigar", Qname, (Score, Cigar), Pos, RefSeq).
    Reads with no sga cluster send nothing; messages come in the reverse of the list order (:74-82)."""
    st, ho, hits, cig = fm.align([q for _, q in qseq_list])
    shift_free = []
    for r in range(len(qseq_list) - 1, -1, -1):
        if st[r] == api.READ_CRASH:
            raise ReferenceCrash("case_clause in bwa:find_seeds for read %r" % (qseq_list[r][0],))
        if st[r] == api.READ_LIMIT:          # not "no clusters": the reference would have aligned this read
            raise LibraryLimit("read %r exceeds this library's limits (1031 bases / 2^22 seed ends)" % (qseq_list[r][0],))
        hs = hits[ho[r]:ho[r + 1]]
        if len(hs) == 0:
            continue
        name, q = qseq_list[r]
        cigars = []
        for h in hs:
            if h["status"] == api.HIT_WINDOW_OOR:
                raise ReferenceCrash("badmatch: window outside the padded reference (worker_bwt.erl:89)")
            if h["status"] != api.HIT_MATCH:
                continue
            ref_len = len(q) + int(h["d"])
            start = int(h["s"]) - ref_len + REF_EXTENSION_LEN
            c = cig[int(h["cigar_off"]):int(h["cigar_off"]) + int(h["cigar_len"])].decode()
            cigars.append(((int(h["score"]), c), int(h["pos"]), padded_ref[start:start + ref_len]))
        if not cigars:
            shift_free.append("no_cigar")
        elif len(cigars) == 1:
            shift_free.append(("cigar", name) + cigars[0])
        else:
            shift_free.append(("cigar", name) + otp_sort_top(cigars))
    return shift_free
