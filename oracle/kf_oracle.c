/*
 * kf_oracle.c -- plain-C CPU restatement of kf2vec's k-mer frequency path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library.  The product never links it.
 *
 * Follows (reference checkout paths):
 *   kf2vec/main.py:308-311  jellyfish count -m k -s 100M -t p -C <file>
 *   kf2vec/main.py:317-319  jellyfish dump -c
 *   kf2vec/main.py:327-328  left merge on the sorted canonical vocabulary, fillna(0)
 *   kf2vec/main.py:332-342  +0.5 pseudocount, fp64 normalisation
 * Jellyfish itself is an external dependency (kmer-jellyfish=1.1.12, kf2vec_env.yml:35) that is not
 * vendored; its behaviour for "count -C" is restated: type sniffed from the first byte, headers
 * skipped, '\n' removed, a break between records, A/C/G/T (either case) -> 0/1/2/3, any other byte
 * resets the window, bin = min(forward, reverse complement).  It rolls the forward and the
 * reverse-complement mer per base the way Jellyfish does (the CUDA path does not: it counts forward
 * mers and folds at the end), so agreement between the two is a real cross-check.
 *
 * Pinning: checked byte-exact against the reference's 7 reproducible toy_example .kf goldens through
 * tests/test_oracle_golden.py (via the NumPy twin and directly).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define KFO_OK 0
#define KFO_ERR_FORMAT -1
#define KFO_ERR_IO -2
#define KFO_ERR_ARG -3

static int8_t g_code[256];
static int g_code_init = 0;

static void init_codes(void) {
    if (g_code_init) return;
    memset(g_code, -1, sizeof g_code);
    g_code['A'] = g_code['a'] = 0;
    g_code['C'] = g_code['c'] = 1;
    g_code['G'] = g_code['g'] = 2;
    g_code['T'] = g_code['t'] = 3;
    g_code_init = 1;
}

typedef struct {
    int k;
    uint64_t mask;
    uint64_t f, r;
    int filled;
    uint64_t *dense; /* [4^k], indexed by canonical code (NULL in list mode) */
    uint64_t *list;  /* list mode (sparse counting, any k <= 31): every canonical occurrence appended */
    size_t nlist;
} roller_t;

static inline void roll_reset(roller_t *R) { R->filled = 0; }

static inline void roll_push(roller_t *R, unsigned char c) {
    int code = g_code[c];
    if (code < 0) { R->filled = 0; return; }
    R->f = ((R->f << 2) | (uint64_t)code) & R->mask;
    R->r = (R->r >> 2) | ((uint64_t)(3 - code) << (2 * (R->k - 1)));
    if (R->filled < R->k) R->filled++;
    if (R->filled == R->k) {
        const uint64_t c = R->f < R->r ? R->f : R->r;
        if (R->dense) R->dense[c]++;
        else R->list[R->nlist++] = c;
    }
}

static size_t skip_newlines(const uint8_t *d, size_t n, size_t p) {
    while (p < n && d[p] == '\n') p++;
    return p;
}
static size_t ignore_line(const uint8_t *d, size_t n, size_t p) {
    const uint8_t *q = (const uint8_t *)memchr(d + p, '\n', n - p);
    return q ? (size_t)(q - d) + 1 : n;
}

static void walk_fasta(const uint8_t *d, size_t n, roller_t *R) {
    size_t p = ignore_line(d, n, 0); /* first header */
    while (p < n) {
        p = skip_newlines(d, n, p);
        if (p >= n) break;
        if (d[p] == '>') { /* header at line start: record break */
            roll_reset(R);
            p = ignore_line(d, n, p);
            continue;
        }
        while (p < n && d[p] != '\n') roll_push(R, d[p++]);
    }
}

static void walk_fastq(const uint8_t *d, size_t n, roller_t *R) {
    size_t p = ignore_line(d, n, 0); /* first '@' header */
    while (p < n) {
        size_t nseq = 0, nq = 0;
        p = skip_newlines(d, n, p);
        while (p < n && d[p] != '+') {
            while (p < n && d[p] != '\n') { roll_push(R, d[p++]); nseq++; }
            p = skip_newlines(d, n, p);
        }
        roll_reset(R);
        if (p >= n) break;
        p = ignore_line(d, n, p); /* '+' line */
        p = skip_newlines(d, n, p);
        while (p < n && nq < nseq) {
            while (p < n && d[p] != '\n') { p++; nq++; }
            p = skip_newlines(d, n, p);
        }
        p = skip_newlines(d, n, p);
        p = ignore_line(d, n, p); /* next '@' header */
    }
}

static uint64_t revcomp(uint64_t x, int k) {
    uint64_t o = 0;
    for (int i = 0; i < k; i++) { o = (o << 2) | (3 - (x & 3)); x >>= 2; }
    return o;
}

int kfo_vocab_size(int k) {
    if (k < 1 || k > 15) return KFO_ERR_ARG;
    uint64_t nb = 1ull << (2 * k);
    return (int)((nb + ((k % 2 == 0) ? (1ull << k) : 0)) / 2);
}

/* canonical counts in vocabulary (sorted canonical) order; returns number of valid k-mers in *total */
int kfo_count_buffer(const uint8_t *data, size_t n, int k, uint64_t *canon_counts, uint64_t *total) {
    init_codes();
    if (k < 1 || k > 15) return KFO_ERR_ARG;
    if (n == 0) return KFO_ERR_FORMAT;
    uint64_t nb = 1ull << (2 * k);
    roller_t R;
    R.k = k; R.mask = nb - 1; R.f = R.r = 0; R.filled = 0;
    R.list = NULL; R.nlist = 0;
    R.dense = (uint64_t *)calloc(nb, sizeof(uint64_t));
    if (!R.dense) return KFO_ERR_ARG;
    if (data[0] == '>') walk_fasta(data, n, &R);
    else if (data[0] == '@') walk_fastq(data, n, &R);
    else { free(R.dense); return KFO_ERR_FORMAT; }
    uint64_t t = 0; size_t j = 0;
    for (uint64_t x = 0; x < nb; x++) {
        if (x <= revcomp(x, k)) { canon_counts[j++] = R.dense[x]; t += R.dense[x]; }
    }
    if (total) *total = t;
    free(R.dense);
    return KFO_OK;
}

/* Sparse counting for any k <= 31: what `jellyfish count -C` + `jellyfish dump -c` list (kf2vec/main.py:135-145, get_kmers):
 * the OBSERVED canonical k-mers with their counts -- here in ascending code order (A0 C1 G2 T3, first base most
 * significant; Jellyfish lists them in hash order).  codes_out / counts_out hold up to cap entries; *n_distinct gets the
 * number found (the arrays are filled only when it fits), *total the number of valid k-mers. */
static int cmp_u64(const void *a, const void *b) {
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
int kfo_count_sparse(const uint8_t *data, size_t n, int k, uint64_t *codes_out, uint64_t *counts_out, size_t cap,
                     uint64_t *n_distinct, uint64_t *total) {
    init_codes();
    if (k < 1 || k > 31) return KFO_ERR_ARG;
    if (n == 0) return KFO_ERR_FORMAT;
    roller_t R;
    R.k = k; R.mask = (1ull << (2 * k)) - 1; R.f = R.r = 0; R.filled = 0;
    R.dense = NULL; R.nlist = 0;
    R.list = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    if (!R.list) return KFO_ERR_ARG;
    if (data[0] == '>') walk_fasta(data, n, &R);
    else if (data[0] == '@') walk_fastq(data, n, &R);
    else { free(R.list); return KFO_ERR_FORMAT; }
    qsort(R.list, R.nlist, sizeof(uint64_t), cmp_u64);
    uint64_t nd = 0;
    for (size_t i = 0; i < R.nlist;) {
        size_t j = i + 1;
        while (j < R.nlist && R.list[j] == R.list[i]) j++;
        if (nd < cap) { codes_out[nd] = R.list[i]; counts_out[nd] = (uint64_t)(j - i); }
        nd++;
        i = j;
    }
    if (n_distinct) *n_distinct = nd;
    if (total) *total = (uint64_t)R.nlist;
    free(R.list);
    return KFO_OK;
}

int kfo_frequencies(const uint64_t *canon_counts, int V, int pseudocount, int raw_cnt, double *out) {
    double sum = 0.0;
    for (int i = 0; i < V; i++) { out[i] = (double)canon_counts[i] + (pseudocount ? 0.5 : 0.0); sum += out[i]; }
    if (!raw_cnt) for (int i = 0; i < V; i++) out[i] = out[i] / sum;
    return KFO_OK;
}

/* ---- multi-threaded driver over in-memory buffers: the CPU baseline ("jellyfish -t p" stand-in) ---- */
typedef struct {
    const uint8_t *const *bufs; const size_t *lens; int n; int k; int V;
    uint64_t *counts; double *freq; int *status; int next; pthread_mutex_t mu;
} job_t;

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int i = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (i >= J->n) break;
        uint64_t tot;
        uint64_t *row = J->counts + (size_t)i * J->V;
        J->status[i] = kfo_count_buffer(J->bufs[i], J->lens[i], J->k, row, &tot);
        if (J->status[i] == KFO_OK && J->freq) kfo_frequencies(row, J->V, 0, 0, J->freq + (size_t)i * J->V);
    }
    return NULL;
}

int kfo_count_buffers_mt(const uint8_t *const *bufs, const size_t *lens, int n, int k, int threads,
                         uint64_t *counts, double *freq, int *status) {
    init_codes();
    int V = kfo_vocab_size(k);
    if (V < 0 || n < 0 || threads < 1) return KFO_ERR_ARG;
    job_t J; J.bufs = bufs; J.lens = lens; J.n = n; J.k = k; J.V = V;
    J.counts = counts; J.freq = freq; J.status = status; J.next = 0;
    pthread_mutex_init(&J.mu, NULL);
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &J);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&J.mu);
    return KFO_OK;
}
