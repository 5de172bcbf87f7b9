"""Error behaviour of the mirror modules: where the reference process crashes, these raise ReferenceCrash."""


class ReferenceCrash(Exception):
    """The reference worker would crash here (case_clause bwa.erl:15-24, badmatch worker_bwt.erl:89, badarg)."""
