/* Declarations of the erl_nif API subset that c_src/bwt_nif.c uses -- for `gcc -fsyntax-only` in an image without OTP.
 * Test infrastructure only; the real header ships with Erlang/OTP (erts/emulator/beam/erl_nif.h). */
#ifndef MOCK_ERL_NIF_H
#define MOCK_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>
typedef uintptr_t ERL_NIF_TERM;
typedef struct enif_environment_t ErlNifEnv;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef struct { size_t size; unsigned char *data; void *ref_bin; void *spare[2]; } ErlNifBinary;
typedef void ErlNifResourceDtor(ErlNifEnv *, void *);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
typedef enum { ERL_NIF_LATIN1 = 1 } ErlNifCharEncoding;
#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2
typedef struct { const char *name; unsigned arity; ERL_NIF_TERM (*fptr)(ErlNifEnv *, int, const ERL_NIF_TERM[]); unsigned flags; } ErlNifFunc;
ErlNifResourceType *enif_open_resource_type(ErlNifEnv *, const char *, const char *, ErlNifResourceDtor *, ErlNifResourceFlags, ErlNifResourceFlags *);
void *enif_alloc_resource(ErlNifResourceType *, size_t);
void enif_release_resource(void *);
ERL_NIF_TERM enif_make_resource(ErlNifEnv *, void *);
int enif_get_resource(ErlNifEnv *, ERL_NIF_TERM, ErlNifResourceType *, void **);
ERL_NIF_TERM enif_make_atom(ErlNifEnv *, const char *);
ERL_NIF_TERM enif_make_int(ErlNifEnv *, int);
ERL_NIF_TERM enif_make_uint(ErlNifEnv *, unsigned);
ERL_NIF_TERM enif_make_int64(ErlNifEnv *, int64_t);
ERL_NIF_TERM enif_make_uint64(ErlNifEnv *, uint64_t);
ERL_NIF_TERM enif_make_string(ErlNifEnv *, const char *, ErlNifCharEncoding);
ERL_NIF_TERM enif_make_string_len(ErlNifEnv *, const char *, size_t, ErlNifCharEncoding);
ERL_NIF_TERM enif_make_tuple(ErlNifEnv *, unsigned, ...);
#define enif_make_tuple2(e, a, b) enif_make_tuple(e, 2, a, b)
#define enif_make_tuple3(e, a, b, c) enif_make_tuple(e, 3, a, b, c)
#define enif_make_tuple4(e, a, b, c, d) enif_make_tuple(e, 4, a, b, c, d)
#define enif_make_tuple6(e, a, b, c, d, f, g) enif_make_tuple(e, 6, a, b, c, d, f, g)
ERL_NIF_TERM enif_make_list(ErlNifEnv *, unsigned, ...);
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_list_from_array(ErlNifEnv *, const ERL_NIF_TERM[], unsigned);
int enif_make_reverse_list(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM *);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv *);
int enif_get_list_length(ErlNifEnv *, ERL_NIF_TERM, unsigned *);
int enif_get_list_cell(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM *, ERL_NIF_TERM *);
int enif_get_tuple(ErlNifEnv *, ERL_NIF_TERM, int *, const ERL_NIF_TERM **);
int enif_get_int(ErlNifEnv *, ERL_NIF_TERM, int *);
int enif_get_string(ErlNifEnv *, ERL_NIF_TERM, char *, unsigned, ErlNifCharEncoding);
int enif_inspect_iolist_as_binary(ErlNifEnv *, ERL_NIF_TERM, ErlNifBinary *);
#define ERL_NIF_INIT(NAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD) \
    const ErlNifFunc *mock_nif_funcs_##NAME(int *n) { *n = (int)(sizeof(FUNCS) / sizeof(FUNCS[0])); (void)LOAD; return FUNCS; }
#endif
