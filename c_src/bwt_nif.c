/*
 * bwt_nif.c -- thin erl_nif shim between the Erlang worker (src/worker_bwt.erl) and libbwt_b200.so.
 *
 * Nothing is computed here: every function marshals Erlang terms to the plain C ABI of include/bwt_b200.h and
 * back.  All calls that touch the GPU are flagged dirty (ERL_NIF_DIRTY_JOB_CPU_BOUND) so that a batch never blocks
 * a normal BEAM scheduler.  The index handle is an Erlang resource: it is immutable after open, so any number of
 * worker processes in one VM may share it (one handle per GPU).
 *
 * Source only in this repository: the image has no OTP (no erl_nif.h).  tests/test_host_logic.py compiles it
 * against tests/mock_erl/erl_nif.h (declarations of the API subset used here) with -fsyntax-only.
 *
 * Erlang side (module bwt_nif, c_src/bwt_nif.erl):
 *   open(FmPath, RefPath, Device)     -> {ok, Index} | {error, {Code, Msg}}          fm_index:get_index/2 + .ref load
 *   info(Index)                       -> {{Pc,Pg,Pt,Last}, Shift, NRows}
 *   find_seeds(Index, [Subseq])       -> [{seeds|no_seeds|too_many_seeds|crash, Sp, Ep, N}]   bwa:find_seeds/6
 *   sga(Index, [Qseq])                -> [[{S,D}] | crash | limit]                   sga:sga/6
 *   sw([{Qseq, Ref}], Device)         -> [no_match | {Score, Cigar}]                 sw:sw/2
 *   align(Index, [Qseq])              -> [[{Status, Score, Cigar, Pos, S, D}] | crash | limit]   worker_bwt.erl:76-97
 *   align_workload(Index, [[Qseq]])   -> [[...as align...]]   a whole {workload, [[Qseq] x 10]} (worker_bwt.erl:64-69) in ONE call
 *   close(Index)                      -> ok        releases the HBM now (dirty scheduler) instead of at garbage collection
 * Per-read atoms: `crash` = the reference worker would crash on this read (case_clause in bwa:find_seeds/6 on a last
 * character outside ACGTN, bwa.erl:15-24); `limit` = THIS library's limit (read longer than 1031 bases or more than
 * 2^22 seed ends), not a crash in the reference: the shim falls back to the Erlang sga/sw for that read, or logs and
 * skips it.
 * Allocation failures come back as {error, enomem}.
 * Qseq / Subseq / Ref are strings (lists of bytes) or binaries.
 */
#include <erl_nif.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bwt_b200.h"

static ErlNifResourceType *INDEX_RES;

typedef struct { bwt_index *h; } index_res;

/* Runs when the last reference to the resource is collected -- on a NORMAL scheduler thread, where cudaFree of tens of
 * gigabytes does not belong: workers call close/1 (dirty) when they stop; this is the fallback for a crashed owner. */
static void index_dtor(ErlNifEnv *env, void *obj) {
    (void)env;
    index_res *r = (index_res *)obj;
    if (r->h) bwt_close(r->h);
    r->h = NULL;
}

static ERL_NIF_TERM mk_enomem(ErlNifEnv *env) {
    return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_atom(env, "enomem"));
}

static ERL_NIF_TERM read_status_atom(ErlNifEnv *env, int32_t rs) {
    return enif_make_atom(env, rs == BWT_READ_LIMIT ? "limit" : "crash");
}

static int on_load(ErlNifEnv *env, void **priv, ERL_NIF_TERM info) {
    (void)priv; (void)info;
    INDEX_RES = enif_open_resource_type(env, NULL, "bwt_index", index_dtor, ERL_NIF_RT_CREATE, NULL);
    return INDEX_RES ? 0 : -1;
}

static ERL_NIF_TERM mk_error(ErlNifEnv *env, int code) {
    const char *msg = bwt_last_error();
    return enif_make_tuple2(env, enif_make_atom(env, "error"),
                            enif_make_tuple2(env, enif_make_int(env, code),
                                             enif_make_string(env, msg ? msg : "", ERL_NIF_LATIN1)));
}

/* a list of iolists -> concatenated bytes + offsets[n+1]; returns 0 on badarg */
typedef struct { uint8_t *data; int64_t *offs; int64_t n; } packed;

static void packed_free(packed *p) { free(p->data); free(p->offs); p->data = NULL; p->offs = NULL; }

static int pack_list(ErlNifEnv *env, ERL_NIF_TERM list, int tuple_field, packed *out) {
    unsigned len;
    if (!enif_get_list_length(env, list, &len)) return 0;
    out->n = (int64_t)len;
    out->offs = (int64_t *)malloc(((size_t)len + 1) * sizeof(int64_t));
    out->data = NULL;
    if (!out->offs) return 0;
    size_t cap = 0, used = 0;
    ERL_NIF_TERM head, tail = list;
    int64_t k = 0;
    out->offs[0] = 0;
    while (enif_get_list_cell(env, tail, &head, &tail)) {
        ERL_NIF_TERM item = head;
        if (tuple_field >= 0) {                       /* element of a {Qseq, Ref} tuple */
            const ERL_NIF_TERM *tup;
            int arity;
            if (!enif_get_tuple(env, head, &arity, &tup) || arity <= tuple_field) { packed_free(out); return 0; }
            item = tup[tuple_field];
        }
        ErlNifBinary bin;
        if (!enif_inspect_iolist_as_binary(env, item, &bin)) { packed_free(out); return 0; }
        if (used + bin.size + 16 > cap) {
            cap = (used + bin.size + 16) * 2;
            uint8_t *grown = (uint8_t *)realloc(out->data, cap);
            if (!grown) { packed_free(out); return 0; }
            out->data = grown;
        }
        memcpy(out->data + used, bin.data, bin.size);
        used += bin.size;
        out->offs[++k] = (int64_t)used;
    }
    if (!out->data) out->data = (uint8_t *)malloc(16);
    if (!out->data) { packed_free(out); return 0; }
    return 1;
}

static int get_index(ErlNifEnv *env, ERL_NIF_TERM t, bwt_index **h) {
    index_res *r;
    if (!enif_get_resource(env, t, INDEX_RES, (void **)&r) || !r->h) return 0;
    *h = r->h;
    return 1;
}

/* open(FmPath, RefPath, Device) */
static ERL_NIF_TERM nif_open(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    char fm[4096], ref[4096];
    int device;
    (void)argc;
    if (enif_get_string(env, argv[0], fm, sizeof fm, ERL_NIF_LATIN1) <= 0 ||
        enif_get_string(env, argv[1], ref, sizeof ref, ERL_NIF_LATIN1) <= 0 || !enif_get_int(env, argv[2], &device))
        return enif_make_badarg(env);
    bwt_index *h = NULL;
    int st = bwt_open(fm, ref, device, &h);
    if (st != BWT_OK) return mk_error(env, st);
    index_res *r = (index_res *)enif_alloc_resource(INDEX_RES, sizeof(index_res));
    r->h = h;
    ERL_NIF_TERM term = enif_make_resource(env, r);
    enif_release_resource(r);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), term);
}

/* info(Index) -> {{Pc,Pg,Pt,Last}, Shift, NRows} : the Meta proplist of fm_index:get_index/2 */
static ERL_NIF_TERM nif_info(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    bwt_index *h;
    bwt_index_info i;
    (void)argc;
    if (!get_index(env, argv[0], &h)) return enif_make_badarg(env);
    int st = bwt_info(h, &i);
    if (st != BWT_OK) return mk_error(env, st);
    return enif_make_tuple3(env,
                            enif_make_tuple4(env, enif_make_uint(env, i.pointers[0]), enif_make_uint(env, i.pointers[1]),
                                             enif_make_uint(env, i.pointers[2]), enif_make_uint(env, i.pointers[3])),
                            enif_make_int64(env, i.shift), enif_make_uint64(env, i.n_rows));
}

/* find_seeds(Index, [Subseq]) */
static ERL_NIF_TERM nif_find_seeds(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    bwt_index *h;
    packed p;
    (void)argc;
    if (!get_index(env, argv[0], &h) || !pack_list(env, argv[1], -1, &p)) return enif_make_badarg(env);
    int64_t n = p.n;
    int32_t *status = (int32_t *)malloc((size_t)(n + 1) * 4), *nm = (int32_t *)malloc((size_t)(n + 1) * 4);
    uint32_t *sp = (uint32_t *)malloc((size_t)(n + 1) * 4), *ep = (uint32_t *)malloc((size_t)(n + 1) * 4);
    ERL_NIF_TERM *items = (ERL_NIF_TERM *)malloc((size_t)(n + 1) * sizeof(ERL_NIF_TERM));
    if (!status || !nm || !sp || !ep || !items) {
        free(status); free(nm); free(sp); free(ep); free(items); packed_free(&p);
        return mk_enomem(env);
    }
    int st = bwt_search_batch(h, p.data, p.offs, n, status, sp, ep, nm);
    ERL_NIF_TERM res;
    if (st != BWT_OK) res = mk_error(env, st);
    else {
        static const char *names[4] = {"seeds", "no_seeds", "too_many_seeds", "crash"};
        for (int64_t k = 0; k < n; k++)
            items[k] = enif_make_tuple4(env, enif_make_atom(env, names[status[k] & 3]), enif_make_uint(env, sp[k]),
                                        enif_make_uint(env, ep[k]), enif_make_int(env, nm[k]));
        res = enif_make_list_from_array(env, items, (unsigned)n);
    }
    free(status); free(nm); free(sp); free(ep); free(items);
    packed_free(&p);
    return res;
}

/* sga(Index, [Qseq]) -> [[{S,D}] | crash] */
static ERL_NIF_TERM nif_sga(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    bwt_index *h;
    packed p;
    (void)argc;
    if (!get_index(env, argv[0], &h) || !pack_list(env, argv[1], -1, &p)) return enif_make_badarg(env);
    int64_t n = p.n, cap = 4 * n + 16, total = 0;
    int32_t *rs = (int32_t *)malloc((size_t)(n + 1) * 4);
    int64_t *clo = (int64_t *)malloc((size_t)(n + 1) * 8), *S = NULL;
    int32_t *D = NULL;
    ERL_NIF_TERM *items = (ERL_NIF_TERM *)malloc((size_t)(n + 1) * sizeof(ERL_NIF_TERM));
    int st = rs && clo && items ? BWT_OK : BWT_ERR_NOMEM;
    while (st == BWT_OK) {
        int64_t *S2 = (int64_t *)realloc(S, (size_t)cap * 8);
        int32_t *D2 = (int32_t *)realloc(D, (size_t)cap * 4);
        if (S2) S = S2;
        if (D2) D = D2;
        if (!S2 || !D2) { st = BWT_ERR_NOMEM; break; }
        st = bwt_sga_batch(h, p.data, p.offs, n, rs, clo, S, D, cap, &total);
        if (st != BWT_ERR_NOSPACE) break;
        cap = total + 16;
        st = BWT_OK;
    }
    ERL_NIF_TERM res;
    if (st == BWT_ERR_NOMEM) res = mk_enomem(env);
    else if (st != BWT_OK) res = mk_error(env, st);
    else {
        for (int64_t r = 0; r < n; r++) {
            if (rs[r] != BWT_READ_OK) { items[r] = read_status_atom(env, rs[r]); continue; }
            ERL_NIF_TERM l = enif_make_list(env, 0);
            for (int64_t k = clo[r + 1] - 1; k >= clo[r]; k--)      /* cons from the tail: list keeps sga's order */
                l = enif_make_list_cell(env, enif_make_tuple2(env, enif_make_int64(env, S[k]), enif_make_int(env, D[k])), l);
            items[r] = l;
        }
        res = enif_make_list_from_array(env, items, (unsigned)n);
    }
    free(rs); free(clo); free(S); free(D); free(items);
    packed_free(&p);
    return res;
}

/* fs_find_seeds(Index, [Qsec], Len, Defects) -> [short | limit | [N]]: what fs:find_seeds(Ref_Seq) returns after
 * seeds:generate_fs(Qsec, Len, Defects) (seeds.erl:11-62), Ref_Seq = the text of the partition */
static ERL_NIF_TERM nif_fs_find_seeds(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    bwt_index *h;
    packed p;
    int len, defects;
    (void)argc;
    if (!get_index(env, argv[0], &h) || !enif_get_int(env, argv[2], &len) || !enif_get_int(env, argv[3], &defects) ||
        !pack_list(env, argv[1], -1, &p))
        return enif_make_badarg(env);
    int64_t n = p.n, cap = 4 * n + 16, total = 0;
    int32_t *qs = (int32_t *)malloc((size_t)(n + 1) * 4);
    int64_t *so = (int64_t *)malloc((size_t)(n + 1) * 8), *seeds = NULL;
    ERL_NIF_TERM *items = (ERL_NIF_TERM *)malloc((size_t)(n + 1) * sizeof(ERL_NIF_TERM));
    int st = qs && so && items ? BWT_OK : BWT_ERR_NOMEM;
    while (st == BWT_OK) {
        int64_t *s2 = (int64_t *)realloc(seeds, (size_t)cap * 8);
        if (!s2) { st = BWT_ERR_NOMEM; break; }
        seeds = s2;
        st = bwt_fs_find_seeds_batch(h, p.data, p.offs, n, len, defects, 0, qs, so, seeds, cap, &total, NULL);
        if (st != BWT_ERR_NOSPACE) break;
        cap = total + 16;
        st = BWT_OK;
    }
    ERL_NIF_TERM res;
    if (st == BWT_ERR_NOMEM) res = mk_enomem(env);
    else if (st != BWT_OK) res = mk_error(env, st);
    else {
        for (int64_t r = 0; r < n; r++) {
            if (qs[r] == BWT_FS_SHORT) { items[r] = enif_make_atom(env, "short"); continue; }
            if (qs[r] == BWT_FS_LIMIT) { items[r] = enif_make_atom(env, "limit"); continue; }
            ERL_NIF_TERM l = enif_make_list(env, 0);
            for (int64_t k = so[r + 1] - 1; k >= so[r]; k--)        /* cons from the tail: descending N, as the accumulator */
                l = enif_make_list_cell(env, enif_make_int64(env, seeds[k]), l);
            items[r] = l;
        }
        res = enif_make_list_from_array(env, items, (unsigned)n);
    }
    free(qs); free(so); free(seeds); free(items);
    packed_free(&p);
    return res;
}

/* sw([{Qseq, Ref}], Device) -> [no_match | {Score, Cigar}] */
static ERL_NIF_TERM nif_sw(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    packed q, r;
    int device;
    (void)argc;
    if (!enif_get_int(env, argv[1], &device) || !pack_list(env, argv[0], 0, &q)) return enif_make_badarg(env);
    if (!pack_list(env, argv[0], 1, &r)) { packed_free(&q); return enif_make_badarg(env); }
    int64_t n = q.n, ccap = 64 * n + 256, ctot = 0;
    int32_t *status = (int32_t *)malloc((size_t)(n + 1) * 4), *score = (int32_t *)malloc((size_t)(n + 1) * 4);
    int32_t *clen = (int32_t *)malloc((size_t)(n + 1) * 4);
    int64_t *coff = (int64_t *)malloc((size_t)(n + 1) * 8);
    char *cig = NULL;
    ERL_NIF_TERM *items = (ERL_NIF_TERM *)malloc((size_t)(n + 1) * sizeof(ERL_NIF_TERM));
    int st = status && score && clen && coff && items ? BWT_OK : BWT_ERR_NOMEM;
    while (st == BWT_OK) {
        char *c2 = (char *)realloc(cig, (size_t)ccap);
        if (!c2) { st = BWT_ERR_NOMEM; break; }
        cig = c2;
        st = bwt_sw_batch(device, q.data, q.offs, r.data, r.offs, n, status, score, coff, clen, cig, ccap, &ctot);
        if (st != BWT_ERR_NOSPACE) break;
        ccap = ctot + 16;
        st = BWT_OK;
    }
    ERL_NIF_TERM res;
    if (st == BWT_ERR_NOMEM) res = mk_enomem(env);
    else if (st != BWT_OK) res = mk_error(env, st);
    else {
        for (int64_t k = 0; k < n; k++)
            items[k] = status[k] == BWT_HIT_MATCH
                           ? enif_make_tuple2(env, enif_make_int(env, score[k]),
                                              enif_make_string_len(env, cig + coff[k], (size_t)clen[k], ERL_NIF_LATIN1))
                           : enif_make_atom(env, "no_match");
        res = enif_make_list_from_array(env, items, (unsigned)n);
    }
    free(status); free(score); free(clen); free(coff); free(cig); free(items);
    packed_free(&q); packed_free(&r);
    return res;
}

/* the worker's inner loop for the reads of p (one bwt_align_batch call) -> items[r] = crash | limit | [{Status, Score,
 * Cigar, Pos, S, D}] in Cigars order (ascending position); 0 = ok, else a bwt_status */
static int align_core(ErlNifEnv *env, bwt_index *h, const packed *p, ERL_NIF_TERM *items) {
    int64_t n = p->n, hcap = 4 * n + 16, ccap = 64 * n + 256, nh = 0, nc = 0;
    int32_t *rs = (int32_t *)malloc((size_t)(n + 1) * 4);
    int64_t *ho = (int64_t *)malloc((size_t)(n + 1) * 8);
    bwt_hit *hits = NULL;
    char *cig = NULL;
    int st = rs && ho ? BWT_OK : BWT_ERR_NOMEM;
    while (st == BWT_OK) {
        bwt_hit *h2 = (bwt_hit *)realloc(hits, (size_t)hcap * sizeof(bwt_hit));
        char *c2 = (char *)realloc(cig, (size_t)ccap);
        if (h2) hits = h2;
        if (c2) cig = c2;
        if (!h2 || !c2) { st = BWT_ERR_NOMEM; break; }
        st = bwt_align_batch(h, p->data, p->offs, n, rs, ho, hits, hcap, &nh, cig, ccap, &nc);
        if (st != BWT_ERR_NOSPACE) break;
        hcap = nh + 16; ccap = nc + 16;
        st = BWT_OK;
    }
    if (st == BWT_OK) {
        static const char *names[3] = {"match", "no_match", "window_oor"};
        for (int64_t r = 0; r < n; r++) {
            if (rs[r] != BWT_READ_OK) { items[r] = read_status_atom(env, rs[r]); continue; }
            ERL_NIF_TERM l = enif_make_list(env, 0);
            for (int64_t k = ho[r + 1] - 1; k >= ho[r]; k--) {
                const bwt_hit *x = &hits[k];
                int s = x->status < 0 || x->status > 2 ? 1 : x->status;
                ERL_NIF_TERM t = enif_make_tuple6(
                    env, enif_make_atom(env, names[s]), enif_make_int(env, x->score),
                    enif_make_string_len(env, cig + x->cigar_off, (size_t)x->cigar_len, ERL_NIF_LATIN1),
                    enif_make_int64(env, x->pos), enif_make_int64(env, x->s), enif_make_int(env, x->d));
                l = enif_make_list_cell(env, t, l);
            }
            items[r] = l;
        }
    }
    free(rs); free(ho); free(hits); free(cig);
    return st;
}

/* align(Index, [Qseq]) -> per read: crash | limit | [{Status, Score, Cigar, Pos, S, D}];
 * Status: match | no_match | window_oor (the reference badmatches there, worker_bwt.erl:89) */
static ERL_NIF_TERM nif_align(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    bwt_index *h;
    packed p;
    (void)argc;
    if (!get_index(env, argv[0], &h) || !pack_list(env, argv[1], -1, &p)) return enif_make_badarg(env);
    ERL_NIF_TERM *items = (ERL_NIF_TERM *)malloc((size_t)(p.n + 1) * sizeof(ERL_NIF_TERM));
    int st = items ? align_core(env, h, &p, items) : BWT_ERR_NOMEM;
    ERL_NIF_TERM res = st == BWT_OK ? enif_make_list_from_array(env, items, (unsigned)p.n)
                                    : st == BWT_ERR_NOMEM ? mk_enomem(env) : mk_error(env, st);
    free(items);
    packed_free(&p);
    return res;
}

/* align_workload(Index, [[Qseq]]) -> [[per read as align/2]]: the whole {workload, [QseqList x 10]} of worker_bwt.erl:64-69
 * in ONE call (2000 reads instead of ten calls of 200: the GPU path is launch-latency bound at that size) */
static ERL_NIF_TERM nif_align_workload(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    bwt_index *h;
    unsigned nlists;
    (void)argc;
    if (!get_index(env, argv[0], &h) || !enif_get_list_length(env, argv[1], &nlists)) return enif_make_badarg(env);
    /* flatten: one packed batch, the list boundaries kept */
    unsigned *cnt = (unsigned *)malloc(((size_t)nlists + 1) * sizeof(unsigned));
    if (!cnt) return mk_enomem(env);
    ERL_NIF_TERM flat = enif_make_list(env, 0), head, tail = argv[1];
    unsigned k = 0;
    /* build the flat list in reverse, then reverse once */
    while (enif_get_list_cell(env, tail, &head, &tail)) {
        unsigned len;
        ERL_NIF_TERM h2, t2 = head;
        if (!enif_get_list_length(env, head, &len)) { free(cnt); return enif_make_badarg(env); }
        cnt[k++] = len;
        while (enif_get_list_cell(env, t2, &h2, &t2)) flat = enif_make_list_cell(env, h2, flat);
    }
    ERL_NIF_TERM fwd;
    if (!enif_make_reverse_list(env, flat, &fwd)) { free(cnt); return enif_make_badarg(env); }
    packed p;
    if (!pack_list(env, fwd, -1, &p)) { free(cnt); return enif_make_badarg(env); }
    ERL_NIF_TERM *items = (ERL_NIF_TERM *)malloc((size_t)(p.n + 1) * sizeof(ERL_NIF_TERM));
    ERL_NIF_TERM *lists = (ERL_NIF_TERM *)malloc(((size_t)nlists + 1) * sizeof(ERL_NIF_TERM));
    int st = items && lists ? align_core(env, h, &p, items) : BWT_ERR_NOMEM;
    ERL_NIF_TERM res;
    if (st == BWT_OK) {
        int64_t at = 0;
        for (unsigned l = 0; l < nlists; l++) { lists[l] = enif_make_list_from_array(env, items + at, cnt[l]); at += cnt[l]; }
        res = enif_make_list_from_array(env, lists, nlists);
    } else res = st == BWT_ERR_NOMEM ? mk_enomem(env) : mk_error(env, st);
    free(items); free(lists); free(cnt);
    packed_free(&p);
    return res;
}

/* close(Index) -> ok: release the device memory now, on a dirty scheduler */
static ERL_NIF_TERM nif_close(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    index_res *r;
    (void)argc;
    if (!enif_get_resource(env, argv[0], INDEX_RES, (void **)&r)) return enif_make_badarg(env);
    if (r->h) { bwt_close(r->h); r->h = NULL; }
    return enif_make_atom(env, "ok");
}

static ErlNifFunc nif_funcs[] = {
    {"open", 3, nif_open, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"info", 1, nif_info, 0},
    {"find_seeds", 2, nif_find_seeds, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"sga", 2, nif_sga, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"fs_find_seeds", 4, nif_fs_find_seeds, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"sw", 2, nif_sw, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"align", 2, nif_align, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"align_workload", 2, nif_align_workload, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"close", 1, nif_close, ERL_NIF_DIRTY_JOB_IO_BOUND},
};

ERL_NIF_INIT(bwt_nif, nif_funcs, on_load, NULL, NULL, NULL)
