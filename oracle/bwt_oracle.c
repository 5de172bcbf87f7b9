/*
 * CPU port of the cloudozer/BWT read-alignment hot path  --  TEST INFRASTRUCTURE ONLY.
 *
 * This is the fast half of the oracle (the slow, literal half is oracle/bwt_oracle.py, and
 * tests/test_oracle_*.py hold the two against each other and against the golden vectors).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (bwt_b200/csrc) never links or calls it.
 *
 * Parity status: the index it consumes is pinned by the shipped GL000193.1_p1.fm; search / vote /
 * SW outputs are pinned only by clause-for-clause fidelity to the Erlang source (no golden outputs
 * exist in the reference tree) -- see oracle/bwt_oracle.py header.
 *
 * Citations are /root/reference file:line.  Rows are 1-based like Erlang's element/2.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* include/bwt.hrl:1-20 */
#define TOLERANCE 20
#define MIN_LEN 8
#define MAX_RANGE 45
#define MAX_LEN 15
#define MATCH 2
#define UNKNOWN 1.6
#define MISMATCH (-1)
#define GAP_PENALTY (-2)
#define GAP_EXT_PENALTY (-1)
#define REF_EXTENSION_LEN 200 /* worker_bwt.erl:17 */

enum { ST_SEEDS = 0, ST_NO_SEEDS = 1, ST_TOO_MANY = 2, ST_CRASH = 3 };
enum { D_LEFT = 0, D_UP = 1, D_UL = 2, D_UNDEF = 3 }; /* Erlang atom order left < up < up_left */

#define OCC_BLK 64

typedef struct {
    uint64_t n;          /* rows, including the $ row */
    const uint8_t *L;    /* ASCII, n bytes (borrowed) */
    const uint32_t *SA;  /* n entries (borrowed) */
    uint32_t *I;         /* literal mode: the reference's stored LF target, 0-based (owned) */
    uint32_t *occ;       /* closed-form mode: counts of A,C,G,T before each 64-row block (owned) */
    uint64_t P[4];       /* 1-based first row of A (=2), C, G, T */
    uint64_t last;
    uint64_t dollar;     /* 1-based row whose L is '$' */
    int64_t shift;
    const uint8_t *ref;  /* padded with 200 'N' both sides (owned copy) */
    uint64_t ref_len;    /* padded length */
    uint8_t *ref_own;
} orc_index;

static int code_of(uint8_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

/* mode 0: literal (I array + linear scans, fmi.erl:25-39).  mode 1: closed form. */
orc_index *orc_index_new(const uint8_t *L, const uint32_t *SA, uint64_t n, const uint32_t pointers[4],
                         int64_t shift, const uint8_t *ref, uint64_t ref_len, int mode) {
    orc_index *x = (orc_index *)calloc(1, sizeof(orc_index));
    x->n = n; x->L = L; x->SA = SA; x->shift = shift;
    x->P[0] = 2; x->P[1] = pointers[0]; x->P[2] = pointers[1]; x->P[3] = pointers[2]; x->last = pointers[3];
    for (uint64_t k = 0; k < n; k++) if (L[k] == '$') { x->dollar = k + 1; break; }
    if (mode == 0) {
        /* bwt.erl:105-120 + fmi.erl:43-62: the k-th row whose L is c gets the 0-based row number of
           the k-th row whose F is c; F-rows of c start at P[c] (1-based) => 0-based P[c]-1. */
        x->I = (uint32_t *)malloc(n * sizeof(uint32_t));
        uint64_t next[4] = { x->P[0] - 1, x->P[1] - 1, x->P[2] - 1, x->P[3] - 1 };
        for (uint64_t k = 0; k < n; k++) {
            int c = code_of(L[k]);
            x->I[k] = (c < 0) ? 0u : (uint32_t)(next[c]++);
        }
    } else {
        uint64_t nb = n / OCC_BLK + 1;
        x->occ = (uint32_t *)malloc(nb * 4 * sizeof(uint32_t));
        uint32_t cnt[4] = {0, 0, 0, 0};
        for (uint64_t k = 0; k < n; k++) {
            if (k % OCC_BLK == 0) memcpy(x->occ + 4 * (k / OCC_BLK), cnt, sizeof cnt);
            int c = code_of(L[k]);
            if (c >= 0) cnt[c]++;
        }
        if (n % OCC_BLK == 0) memcpy(x->occ + 4 * (n / OCC_BLK), cnt, sizeof cnt);
    }
    x->ref_len = ref_len + 2 * REF_EXTENSION_LEN;          /* worker_bwt.erl:53-55 */
    x->ref_own = (uint8_t *)malloc(x->ref_len);
    memset(x->ref_own, 'N', x->ref_len);
    memcpy(x->ref_own + REF_EXTENSION_LEN, ref, ref_len);
    x->ref = x->ref_own;
    return x;
}

void orc_index_free(orc_index *x) {
    if (!x) return;
    free(x->I); free(x->occ); free(x->ref_own); free(x);
}

/* occ(c, e) = #{ rows e' <= e : L[e'] == c }, e 1-based, e may be 0 */
static uint64_t occ_cf(const orc_index *x, int c, uint64_t e) {
    uint64_t b = e / OCC_BLK, r = x->occ[4 * b + c];
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t k = b * OCC_BLK; k < e; k++) r += (x->L[k] == (uint8_t)sym[c]);
    return r;
}

/* One bwa.erl:36-43 step.  Returns 1 and the new interval if se_down found the character. */
static int lf_step(const orc_index *x, uint64_t sp, uint64_t ep, uint8_t ch, uint64_t *sp1, uint64_t *ep1) {
    if (x->I) {
        uint64_t e = sp;                                    /* fmi.erl:25-31 */
        for (;;) {
            if (e == ep) return 0;
            if (x->L[e - 1] == ch) break;
            e++;
        }
        *sp1 = x->I[e - 1];
        e = ep;                                             /* fmi.erl:35-39 */
        while (x->L[e - 1] != ch) e--;
        *ep1 = x->I[e - 1];
        return 1;
    }
    int c = code_of(ch);
    if (c < 0 || sp >= ep) return 0;   /* '$' in a read is handled by orc_find_seeds */
    uint64_t a = occ_cf(x, c, sp - 1), b = occ_cf(x, c, ep - 1);
    if (b <= a) return 0;
    *sp1 = x->P[c] - 1 + a;
    *ep1 = x->P[c] - 2 + occ_cf(x, c, ep);
    return 1;
}

/* bwa.erl:12-45.  sub[0..len) ; out: status, final Sp/Ep (1-based rows) and N (#chars matched). */
int orc_find_seeds(const orc_index *x, const uint8_t *sub, int len, uint64_t *o_sp, uint64_t *o_ep, int *o_n) {
    uint64_t sp, ep;
    switch (sub[len - 1]) {                                 /* bwa.erl:15-24 */
    case 'A': case 'N': sp = 1; ep = x->P[1] - 1; break;
    case 'C': sp = x->P[1]; ep = x->P[2] - 1; break;
    case 'G': sp = x->P[2]; ep = x->P[3] - 1; break;
    case 'T': sp = x->P[3]; ep = x->last; break;
    default: return ST_CRASH;                               /* case_clause */
    }
    int n = 1, k = len - 2, checked;
    for (;;) {
        if (n == MAX_LEN) checked = 1;                      /* bwa.erl:31-32 */
        else if (k >= 0 && sub[k] == 'N') checked = 1;      /* bwa.erl:33-34 */
        else if (k >= 0) {                                  /* bwa.erl:36-43 */
            uint64_t s1, e1;
            if (sub[k] == '$') {
                /* the $ row's I is 0, so a hit moves to (0,0): every later clause ends in
                   get_seed_ends(FM,0,0,N) = no_seeds for N<8, element(0,FM) badarg otherwise */
                if (sp < ep && x->dollar >= sp && x->dollar <= ep - 1) {
                    n++; *o_sp = 0; *o_ep = 0; *o_n = n;
                    return n < MIN_LEN ? ST_NO_SEEDS : ST_CRASH;
                }
                checked = 0; break;
            }
            if (lf_step(x, sp, ep, sub[k], &s1, &e1)) { sp = s1; ep = e1; k--; n++; continue; }
            checked = 0;
        } else checked = 1;                                 /* bwa.erl:44-45 */
        break;
    }
    *o_sp = sp; *o_ep = ep; *o_n = n;
    if (checked && ep - sp > MAX_RANGE) return ST_TOO_MANY;
    return n >= MIN_LEN ? ST_SEEDS : ST_NO_SEEDS;           /* fmi.erl:89-91 */
}

/* sga.erl:69-76: returns -1 for no_more_subseqs, else Skipped; the new length is len - Skipped */
static int skip_ns(const uint8_t *q, int len) {
    if (len < MIN_LEN) return -1;
    int j = 0, skipped = 0, k = len - 1;
    for (;;) {
        if (j == MIN_LEN) return skipped;
        if (k < 0) return -1;
        if (q[k] == 'N') { skipped += j; j = 1; } else j++;
        k--;
    }
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

typedef struct { int64_t *v; size_t n, cap; } vec64;
static void vpush(vec64 *v, int64_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 256; v->v = (int64_t *)realloc(v->v, v->cap * sizeof(int64_t)); }
    v->v[v->n++] = x;
}

/* statistics the bench reports (definitions: SURVEY.md 8d) */
typedef struct { uint64_t searches, lf_steps, located, sw_calls, cells; } orc_stats;

/* sga.erl:17-64.  Writes up to cap {S,D} pairs (descending S); returns the count or -1 on a reference crash.
   seeds (optional) receives every collected seed end (unsorted). */
static int64_t sga_impl(const orc_index *x, const uint8_t *q, int len, int64_t *S, int64_t *D, int64_t cap,
                        vec64 *acc, orc_stats *st) {
    int qty = 0;
    int64_t shift = 0;
    acc->n = 0;
    for (;;) {
        int skip = skip_ns(q, len);
        if (skip < 0) break;
        len -= skip;
        uint64_t sp, ep; int n;
        int stt = orc_find_seeds(x, q, len, &sp, &ep, &n);
        if (st) { st->searches++; st->lf_steps += (uint64_t)(n - 1); }
        if (stt == ST_CRASH) return -1;
        if (stt == ST_SEEDS) {
            for (uint64_t e = sp; e <= ep; e++) vpush(acc, (int64_t)x->SA[e - 1] + n + shift + skip);
            if (st) st->located += ep - sp + 1;
            qty++;
        }
        len -= MIN_LEN;
        shift += MIN_LEN + skip;
    }
    if (qty < 2) return 0;                                   /* sga.erl:42-43 */
    qsort(acc->v, acc->n, sizeof(int64_t), cmp_i64);
    int need = 2 + qty / 7;
    /* emit ascending into a temp, then reverse (the Erlang accumulator is prepended to) */
    int64_t cnt = 0, p1 = acc->v[0], count = 1, dist = 0;
    vec64 tmp = {0, 0, 0};
    for (size_t k = 1; k < acc->n; k++) {
        int64_t p2 = acc->v[k];
        if (p2 - p1 <= TOLERANCE) { count++; dist += p2 - p1; p1 = p2; }
        else { if (count >= need) { vpush(&tmp, p1); vpush(&tmp, dist); } p1 = p2; count = 1; dist = 0; }
    }
    if (count >= need) { vpush(&tmp, p1); vpush(&tmp, dist); }
    cnt = (int64_t)(tmp.n / 2);
    for (int64_t k = 0; k < cnt && k < cap; k++) { S[k] = tmp.v[2 * (cnt - 1 - k)]; D[k] = tmp.v[2 * (cnt - 1 - k) + 1]; }
    free(tmp.v);
    return cnt;
}

int64_t orc_sga(const orc_index *x, const uint8_t *q, int len, int64_t *S, int64_t *D, int64_t cap) {
    vec64 acc = {0, 0, 0};
    int64_t r = sga_impl(x, q, len, S, D, cap, &acc, NULL);
    free(acc.v);
    return r;
}

/* all seed ends of one read, in collection order is irrelevant (they are voted after a sort): returns count */
int64_t orc_seed_ends(const orc_index *x, const uint8_t *q, int len, int64_t *out, int64_t cap) {
    vec64 acc = {0, 0, 0};
    int64_t S, D;
    int64_t r = sga_impl(x, q, len, &S, &D, 0, &acc, NULL);
    int64_t n = (int64_t)acc.n;
    if (r < 0) n = -1;
    else { qsort(acc.v, acc.n, sizeof(int64_t), cmp_i64); for (int64_t k = 0; k < n && k < cap; k++) out[k] = acc.v[k]; }
    free(acc.v);
    return n;
}

/* sw.erl:164-202, returned as value*10 so that 1.6 stays an exact integer 16 for the int/double split */
static int sigma10(uint8_t q, uint8_t r) {
    if (q == 'N' && r == 'N') return 16;
    if (q == r) return 20;
    if (q == 'N' || r == 'N') return 16;
    switch (r) {
    case 'R': return (q == 'A' || q == 'G') ? 20 : -10;
    case 'Y': return (q == 'C' || q == 'T') ? 20 : -10;
    case 'B': return (q == 'A') ? -10 : 20;
    case 'D': return (q == 'C') ? -10 : 20;
    case 'K': return (q == 'G' || q == 'T') ? 20 : -10;
    case 'M': return (q == 'A' || q == 'C') ? 20 : -10;
    case 'S': return (q == 'G' || q == 'C') ? 20 : -10;
    case 'W': return (q == 'A' || q == 'T') ? 20 : -10;
    case 'V': return (q == 'T') ? -10 : 20;
    default: return -10;
    }
}

static int emit_cigar(const uint8_t *dirs, int lq, int lr, int j, char *cigar, int cap) {
    /* sw.erl:116-160 as a plain traceback (equality with the clause machine is tested in Python) */
    int i = lq, nops = 0;
    char *ops = (char *)malloc((size_t)lq + lr + 2);
    while (i > 0) {
        uint8_t d = dirs[(size_t)i * (lr + 1) + j];
        if (d == D_UL) { ops[nops++] = 'M'; i--; j--; }
        else if (d == D_UP) { ops[nops++] = 'I'; i--; }
        else { ops[nops++] = 'D'; j--; }
    }
    int w = 0;
    for (int k = nops - 1; k >= 0;) {
        int m = k;
        while (m >= 0 && ops[m] == ops[k]) m--;
        char buf[24];
        int l = snprintf(buf, sizeof buf, "%d%c", k - m, ops[k]);
        if (w + l < cap) memcpy(cigar + w, buf, (size_t)l);
        w += l;
        k = m;
    }
    if (w < cap) cigar[w] = 0;
    free(ops);
    return w;
}

/* sw:sw/2 (sw.erl:47-160).  Returns 0 = match (score, cigar filled), 1 = no_match.
   force_double != 0 runs the IEEE-double path even for N-free inputs (they must agree).
   *cells gets rows_done * Lr. */
int orc_sw(const uint8_t *q, int lq, const uint8_t *r, int lr, int force_double,
           int64_t *score, char *cigar, int cap, int *cigar_len, uint64_t *cells) {
    if (cells) *cells = 0;
    if (lq <= 0 || lr <= 0) return 1;                        /* sw.erl:54 */
    int has_n = force_double;
    for (int k = 0; k < lq && !has_n; k++) has_n = (q[k] == 'N');
    for (int k = 0; k < lr && !has_n; k++) has_n = (r[k] == 'N');
    uint8_t *dirs = (uint8_t *)malloc((size_t)(lq + 1) * (lr + 1));
    memset(dirs, D_UNDEF, (size_t)lr + 1);
    int W = lr + 1, ret = 1, jstart = 0;
    long thr = -MATCH * (lq / 2);                            /* sw.erl:62 */
    if (!has_n) {
        int *pv = (int *)calloc((size_t)W, sizeof(int)), *cv = (int *)malloc((size_t)W * sizeof(int));
        uint8_t *pd = dirs;
        int vmax = 0, aborted = 0;
        for (int i = 1; i <= lq; i++) {
            if (vmax < thr) { aborted = 1; break; }          /* sw.erl:69-71 */
            uint8_t *cd = dirs + (size_t)i * W;
            cv[0] = pv[0] + (i == 1 ? GAP_PENALTY : GAP_EXT_PENALTY); /* sw.erl:72-77 */
            cd[0] = D_UP;
            int rowmax = cv[0];
            for (int j = 1; j < W; j++) {                    /* sw.erl:88-98 */
                int bv = pv[j - 1] + sigma10(q[i - 1], r[j - 1]) / 10, bd = D_UL;
                int c = cv[j - 1] + (cd[j - 1] == D_LEFT ? GAP_EXT_PENALTY : GAP_PENALTY);
                if (c > bv) { bv = c; bd = D_LEFT; }          /* left wins only when strictly larger */
                c = pv[j] + (pd[j] == D_UP ? GAP_EXT_PENALTY : GAP_PENALTY);
                if (c > bv || (c == bv && bd == D_LEFT)) { bv = c; bd = D_UP; }
                cv[j] = bv; cd[j] = (uint8_t)bd;
                if (bv > rowmax) rowmax = bv;
            }
            if (cells) *cells += (uint64_t)lr;
            int *t = pv; pv = cv; cv = t; pd = cd;
            vmax = rowmax; thr += MATCH;
        }
        if (!aborted && !(vmax < thr)) {
            ret = 0; *score = vmax;
            for (jstart = 0; pv[jstart] != vmax; jstart++) {}  /* sw.erl:110-112 */
        }
        free(pv); free(cv);
    } else {
        double *pv = (double *)calloc((size_t)W, sizeof(double)), *cv = (double *)malloc((size_t)W * sizeof(double));
        uint8_t *pd = dirs;
        double vmax = 0; int aborted = 0;
        for (int i = 1; i <= lq; i++) {
            if (vmax < (double)thr) { aborted = 1; break; }
            uint8_t *cd = dirs + (size_t)i * W;
            cv[0] = pv[0] + (i == 1 ? GAP_PENALTY : GAP_EXT_PENALTY);
            cd[0] = D_UP;
            double rowmax = cv[0];
            for (int j = 1; j < W; j++) {
                int s10 = sigma10(q[i - 1], r[j - 1]);
                double sg = (s10 == 16) ? UNKNOWN : (double)(s10 / 10);
                double bv = pv[j - 1] + sg; int bd = D_UL;
                double c = cv[j - 1] + (cd[j - 1] == D_LEFT ? GAP_EXT_PENALTY : GAP_PENALTY);
                if (c > bv) { bv = c; bd = D_LEFT; }
                c = pv[j] + (pd[j] == D_UP ? GAP_EXT_PENALTY : GAP_PENALTY);
                if (c > bv || (c == bv && bd == D_LEFT)) { bv = c; bd = D_UP; }
                cv[j] = bv; cd[j] = (uint8_t)bd;
                if (bv > rowmax) rowmax = bv;
            }
            if (cells) *cells += (uint64_t)lr;
            double *t = pv; pv = cv; cv = t; pd = cd;
            vmax = rowmax; thr += MATCH;
        }
        if (!aborted && !(vmax < (double)thr)) {
            ret = 0; *score = (int64_t)vmax;                 /* trunc/1, sw.erl:123 */
            for (jstart = 0; pv[jstart] != vmax; jstart++) {}
        }
        free(pv); free(cv);
    }
    if (ret == 0) {
        int l = emit_cigar(dirs, lq, lr, jstart, cigar, cap);
        if (cigar_len) *cigar_len = l;
    }
    free(dirs);
    return ret;
}

/* ------------------------------------------------------------------------------------------------
 * worker_bwt.erl:73-108 for a batch of reads, sharded over threads (contiguous read ranges).
 * Output per read: number of sga clusters, then per cluster (reference order = DESCENDING S):
 *   S, D, status (0 match, 1 no_match, 2 window out of range = badmatch), score, pos, cigar.
 * The caller turns that into the reference's ascending Cigars list.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t S, D, pos, score;
    int32_t status, cigar_len;
    char cigar[96];          /* long CIGARs are truncated here; cigar_len holds the true length */
} orc_cand;

typedef struct {
    const orc_index *x; const uint8_t *reads; const int64_t *off; int64_t lo, hi;
    int64_t *ncand; orc_cand *cand; int64_t cap_per_read; orc_stats st; int with_sw;
} job_t;

static void *job_run(void *arg) {
    job_t *jb = (job_t *)arg;
    vec64 acc = {0, 0, 0};
    int64_t cap = jb->cap_per_read;
    int64_t *S = (int64_t *)malloc((size_t)cap * 8), *D = (int64_t *)malloc((size_t)cap * 8);
    for (int64_t r = jb->lo; r < jb->hi; r++) {
        const uint8_t *q = jb->reads + jb->off[r];
        int lq = (int)(jb->off[r + 1] - jb->off[r]);
        int64_t c = sga_impl(jb->x, q, lq, S, D, cap, &acc, &jb->st);
        jb->ncand[r] = c;
        if (c < 0 || !jb->with_sw) {
            for (int64_t k = 0; k < c && k < cap; k++) { orc_cand *o = jb->cand + r * cap + k; memset(o, 0, sizeof *o); o->S = S[k]; o->D = D[k]; o->status = -1; }
            continue;
        }
        for (int64_t k = 0; k < c && k < cap; k++) {
            orc_cand *o = jb->cand + r * cap + k;
            memset(o, 0, sizeof *o);
            o->S = S[k]; o->D = D[k];
            int64_t ref_len = lq + D[k];                                  /* worker_bwt.erl:86 */
            int64_t start = S[k] - ref_len + REF_EXTENSION_LEN;           /* worker_bwt.erl:87 */
            o->pos = S[k] - ref_len + jb->x->shift;                       /* worker_bwt.erl:94 */
            if (start < 0 || (uint64_t)(start + ref_len) > jb->x->ref_len) { o->status = 2; continue; }
            uint64_t cells = 0;
            o->status = orc_sw(q, lq, jb->x->ref + start, (int)ref_len, 0, &o->score, o->cigar, (int)sizeof o->cigar,
                               &o->cigar_len, &cells);
            jb->st.sw_calls++; jb->st.cells += cells;
        }
    }
    free(S); free(D); free(acc.v);
    return NULL;
}

int orc_align_batch(const orc_index *x, const uint8_t *reads, const int64_t *off, int64_t nreads, int nthreads,
                    int with_sw, int64_t cap_per_read, int64_t *ncand, orc_cand *cand, orc_stats *stats) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; job_t jb[256];
    for (int t = 0; t < nthreads; t++) {
        memset(&jb[t], 0, sizeof(job_t));
        jb[t].x = x; jb[t].reads = reads; jb[t].off = off;
        jb[t].lo = nreads * t / nthreads; jb[t].hi = nreads * (t + 1) / nthreads;
        jb[t].ncand = ncand; jb[t].cand = cand; jb[t].cap_per_read = cap_per_read; jb[t].with_sw = with_sw;
        pthread_create(&th[t], NULL, job_run, &jb[t]);
    }
    orc_stats tot = {0, 0, 0, 0, 0};
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        tot.searches += jb[t].st.searches; tot.lf_steps += jb[t].st.lf_steps; tot.located += jb[t].st.located;
        tot.sw_calls += jb[t].st.sw_calls; tot.cells += jb[t].st.cells;
    }
    if (stats) *stats = tot;
    return 0;
}

size_t orc_sizeof_cand(void) { return sizeof(orc_cand); }
