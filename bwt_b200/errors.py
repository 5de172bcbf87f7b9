"""Error behaviour of the mirror modules: where the reference process crashes, these raise ReferenceCrash."""


class ReferenceCrash(Exception):
    """The reference worker would crash here (case_clause bwa.erl:15-24, badmatch worker_bwt.erl:89, badarg)."""


class LibraryLimit(ValueError):
    """THIS library's limit, not a crash of the reference (a read longer than BWT_MAX_READ_LEN = 1031 bases, or with more
    than 2^22 seed ends): the reference would align the read; the caller falls back or skips it explicitly."""
