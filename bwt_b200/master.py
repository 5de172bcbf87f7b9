"""Mirror of the data-path half of master.erl: workload framing (master.erl:59,162-178) and the result sink
(master.erl:149-158).  The gen_server / navel plumbing of the reference is out of scope."""
from . import worker_bwt
from .fastq import Reader

WORKLOAD_SIZE = 200     # master.erl:59
WATERLINE = 10          # worker_bwt.erl:16
SCORE_FILTER = 265      # master.erl:149


def produce_workload(reader: Reader, n: int, state: dict):
    """master.erl:162-178 -> [[(name, seq)]] newest list first | "stop"."""
    acc = []
    while n > 0:
        if state.get("fastq_eof"):
            break
        r = reader.read_seqs(WORKLOAD_SIZE)
        if r == "eof":
            state["fastq_eof"] = True
            return acc
        acc.insert(0, r[1])
        n -= 1
    if n > 0 and state.get("fastq_eof") and not acc:
        return "stop"
    return acc


def sink_line(msg, chromosome: str):
    """master.erl:149-158: the line the master prints for a cigar message, None when it is dropped."""
    if msg == "no_cigar":
        return None
    _, name, (score, cigar), pos, refseq = msg
    if score < SCORE_FILTER:
        return None
    return "\n%s      %s      %d      %s      %d      %s\n" % (name.decode("latin1"), chromosome, pos, cigar, score,
                                                                refseq.decode("latin1"))


def run(fm, padded_ref: bytes, fastq_text: bytes, chromosome: str):
    """master:run/4 + one worker, in process: every workload list goes through the GPU path; returns
    (printed lines, reads dealt) in the order a single worker would produce them."""
    reader = Reader(fastq_text)
    state = {"fastq_eof": False}
    lines, dealt = [], 0
    while True:
        wl = produce_workload(reader, WATERLINE, state)
        if wl == "stop":
            break
        for qseq_list in wl:
            dealt += len(qseq_list)
            for msg in worker_bwt.process_list(fm, padded_ref, qseq_list):
                ln = sink_line(msg, chromosome)
                if ln is not None:
                    lines.append(ln)
    return lines, dealt
