"""bwt_b200 -- B200-native (sm_100a) read-alignment hot path of cloudozer/BWT behind a C ABI.

Host-side mirror of the reference's Erlang modules (fm_index, fmi, bwa, sga, sw, worker_bwt) over
``libbwt_b200.so`` (bwt_b200/csrc).  There is no CPU fallback: importing the compute modules without
the built CUDA library raises.
"""
__version__ = "0.1.0"
