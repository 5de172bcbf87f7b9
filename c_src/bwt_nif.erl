%% bwt_nif -- Erlang face of libbwt_b200.so (c_src/bwt_nif.c).  Add this file to the reference's src/ and
%% build c_src/bwt_nif.c into priv/bwt_nif.so (see INTEGRATION.md).  Every function is replaced by its NIF at load.
-module(bwt_nif).
-export([open/3, info/1, find_seeds/2, sga/2, fs_find_seeds/4, sw/2, align/2, align_workload/2, close/1]).
-on_load(init/0).

init() ->
    PrivDir = case code:priv_dir(bwt) of {error, _} -> "priv"; D -> D end,
    erlang:load_nif(filename:join(PrivDir, "bwt_nif"), 0).

%% open(FmPath, RefPath, Device) -> {ok, Index} | {error, {Code, Msg}}
open(_FmPath, _RefPath, _Device) -> erlang:nif_error(nif_not_loaded).
%% info(Index) -> {{Pc,Pg,Pt,Last}, Shift, NRows}
info(_Index) -> erlang:nif_error(nif_not_loaded).
%% find_seeds(Index, [Subseq]) -> [{seeds|no_seeds|too_many_seeds|crash, Sp, Ep, N}]
find_seeds(_Index, _Subseqs) -> erlang:nif_error(nif_not_loaded).
%% sga(Index, [Qseq]) -> [[{S,D}] | crash | limit]      (limit: this library's limit, not a crash of the reference)
sga(_Index, _Qseqs) -> erlang:nif_error(nif_not_loaded).
%% fs_find_seeds(Index, [Qsec], Len, Defects) -> [short | limit | [N]]   (seeds:generate_fs/3 + fs:find_seeds/1)
fs_find_seeds(_Index, _Qsecs, _Len, _Defects) -> erlang:nif_error(nif_not_loaded).
%% sw([{Qseq, Ref}], Device) -> [no_match | {Score, Cigar}]
sw(_Pairs, _Device) -> erlang:nif_error(nif_not_loaded).
%% align(Index, [Qseq]) -> [crash | limit | [{match|no_match|window_oor, Score, Cigar, Pos, S, D}]]
align(_Index, _Qseqs) -> erlang:nif_error(nif_not_loaded).
%% align_workload(Index, [[Qseq]]) -> [[as align/2]]: a whole {workload, [QseqList]} (worker_bwt.erl:64-69) in one call
align_workload(_Index, _QseqLists) -> erlang:nif_error(nif_not_loaded).
%% close(Index) -> ok: releases the device memory now (dirty scheduler) instead of at garbage collection
close(_Index) -> erlang:nif_error(nif_not_loaded).
