#!/usr/bin/env escript
%% -*- erlang -*-
%%
%% Runs the UNMODIFIED reference (cloudozer/BWT: src/fmi.erl, bwa.erl, sga.erl, sw.erl) on its own shipped index
%% (fm_indices/GL000193.1_p1.fm / .ref) over the committed input set tests/golden/reference_inputs.txt and writes
%% tests/golden/reference_outputs.json -- the file that pins the oracle (and through it the CUDA path) to the reference's
%% own outputs for rows A2-A14 of SURVEY.md 8(a).  Needs Erlang/OTP (any version from 17 on); the build image has none,
%% so this is a recipe for a box that does:
%%
%%     escript scripts/make_reference_golden.escript /path/to/cloudozer-BWT tests/golden/reference_inputs.txt \
%%             tests/golden/reference_outputs.json
%%
%% tests/test_reference_golden.py consumes the output file when it exists (CPU: oracle vs file; GPU: CUDA path vs file).
%%
%% Input lines (written by scripts/make_reference_inputs.py):
%%     S <subsequence>        -> bwa:find_seeds/6                      (bwa.erl:12-45, fmi.erl:25-39,89-100)
%%     R <read>               -> sga:sga/6 + the worker's fold and top-1 pick, copied from worker_bwt.erl:84-106
%%     P <query> <window>     -> sw:sw/2                               (sw.erl:47-160)
%% The only code here that is not the reference's is the driver: the reference's logging dependency `lager` (fetched from
%% the network by rebar) is not needed by the four modules and is not compiled.

main([RefRoot, In, Out]) ->
    Tmp = filename:join(case os:getenv("TMPDIR") of false -> "/tmp"; T -> T end, "bwt_reference_beam"),
    ok = filelib:ensure_dir(filename:join(Tmp, "x")),
    Inc = filename:join(RefRoot, "include"),
    lists:foreach(fun(M) ->
                          Src = filename:join([RefRoot, "src", M ++ ".erl"]),
                          case compile:file(Src, [{i, Inc}, {outdir, Tmp}, report_errors]) of
                              {ok, _} -> ok;
                              Err -> io:format(standard_error, "cannot compile ~s: ~p~n", [Src, Err]), halt(1)
                          end
                  end, ["fmi", "bwa", "sga", "sw"]),
    true = code:add_patha(Tmp),
    {ok, Bin} = file:read_file(filename:join([RefRoot, "fm_indices", "GL000193.1_p1.fm"])),
    {Meta, FM} = binary_to_term(Bin),                                         % fm_index.erl:154-156
    {Pc, Pg, Pt, Last} = proplists:get_value(pointers, Meta),
    Shift = proplists:get_value(shift, Meta),
    {ok, Ref0} = file:read_file(filename:join([RefRoot, "fm_indices", "GL000193.1_p1.ref"])),
    Ext = list_to_binary(lists:duplicate(200, $N)),                            % worker_bwt.erl:17,53-55
    Ref = <<Ext/binary, Ref0/binary, Ext/binary>>,
    {ok, Data} = file:read_file(In),
    Lines = [L || L <- binary:split(Data, <<"\n">>, [global]), L =/= <<>>],
    Ctx = {FM, Pc, Pg, Pt, Last, Shift, Ref},
    Items = [handle(binary:split(L, <<" ">>, [global]), Ctx) || L <- Lines],
    Head = io_lib:format("{\"otp_release\": \"~s\", \"reference\": \"cloudozer/BWT src/{fmi,bwa,sga,sw}.erl unmodified; worker fold copied "
                         "from worker_bwt.erl:84-106\", \"index\": \"fm_indices/GL000193.1_p1\", \"items\": [~n", [erlang:system_info(otp_release)]),
    ok = file:write_file(Out, [Head, join(Items, ",\n"), "\n]}\n"]),
    io:format("~b items -> ~s~n", [length(Items), Out]);
main(_) ->
    io:format(standard_error, "usage: make_reference_golden.escript <reference root> <inputs.txt> <outputs.json>~n", []),
    halt(2).

join([], _) -> [];
join([X], _) -> [X];
join([X | T], Sep) -> [X, Sep | join(T, Sep)].

%% -- bwa:find_seeds/6 ----------------------------------------------------------------------------------------------
handle([<<"S">>, Sub], {FM, Pc, Pg, Pt, Last, _, _}) ->
    Res = try bwa:find_seeds(FM, Pc, Pg, Pt, Last, binary_to_list(Sub)) of
              no_seeds -> "\"no_seeds\"";
              too_many_seeds -> "\"too_many_seeds\"";
              L when is_list(L) -> ints(L)
          catch _:_ -> "\"crash\""
          end,
    io_lib:format("{\"kind\": \"S\", \"in\": \"~s\", \"out\": ~s}", [Sub, Res]);

%% -- sga:sga/6, then worker_bwt.erl:84-106 word for word (minus the navel casts) ------------------------------------
handle([<<"R">>, Read], {FM, Pc, Pg, Pt, Last, Shift, Ref}) ->
    Qsec = binary_to_list(Read),
    Out = try sga:sga(FM, Pc, Pg, Pt, Last, Qsec) of
              Seeds1 ->
                  Folded = try
                               Cigars = lists:foldl(fun({S, D}, Acc) ->
                                                            Ref_len = length(Qsec) + D,
                                                            Start_pos = S - Ref_len + 200,
                                                            <<_:Start_pos/bytes, Ref_seq:Ref_len/bytes, _/binary>> = Ref,
                                                            Ref_seq1 = binary_to_list(Ref_seq),
                                                            case sw:sw(Qsec, Ref_seq1) of
                                                                no_match -> Acc;
                                                                Cigar -> [{Cigar, S - Ref_len + Shift, Ref_seq1} | Acc]
                                                            end
                                                    end, [], Seeds1),
                               Top = case Cigars of
                                         [] -> "\"no_cigar\"";
                                         [{C1, P1, _}] -> cigar(C1, P1);
                                         _ ->
                                             [{TopCigar, P, _} | _] = lists:sort(fun({{R1, _}, _, _}, {{R2, _}, _, _}) -> R1 > R2 end, Cigars),
                                             cigar(TopCigar, P)
                                     end,
                               io_lib:format("\"cigars\": [~s], \"top\": ~s", [join([cigar(C, P) || {C, P, _} <- Cigars], ", "), Top])
                           catch _:_ -> "\"cigars\": \"badmatch\", \"top\": \"badmatch\""
                           end,
                  io_lib:format("\"sga\": [~s], ~s", [join([io_lib:format("[~b, ~b]", [S, D]) || {S, D} <- Seeds1], ", "), Folded])
          catch _:_ -> "\"sga\": \"crash\""
          end,
    io_lib:format("{\"kind\": \"R\", \"in\": \"~s\", ~s}", [Read, Out]);

%% -- sw:sw/2 ---------------------------------------------------------------------------------------------------------
handle([<<"P">>, Q, W], _) ->
    Res = try sw:sw(binary_to_list(Q), binary_to_list(W)) of
              no_match -> "\"no_match\"";
              {Score, Cigar} -> io_lib:format("[~b, \"~s\"]", [Score, Cigar])
          catch _:_ -> "\"crash\""
          end,
    io_lib:format("{\"kind\": \"P\", \"in\": [\"~s\", \"~s\"], \"out\": ~s}", [Q, W, Res]).

cigar({Score, Cigar}, Pos) -> io_lib:format("[~b, \"~s\", ~b]", [Score, Cigar, Pos]).
ints(L) -> ["[", join([integer_to_list(X) || X <- L], ", "), "]"].
