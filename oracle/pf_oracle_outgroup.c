/*
 * pf_oracle_outgroup.c -- CPU restatement of primerForge's outgroup screen
 * (/root/reference/bin/removeOutgroupPrimers.py).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under primerforge_b200/ may import, link or execute this file; it is the
 * checker used by tests/ and by bench.py's cpu_baseline leg of the outgroup block.  It follows the REFERENCE's
 * algorithm on STRINGS -- every contig forward and reverse-complemented, byte comparison of every window with the
 * primer strings, per-contig per-strand lists of starts, the nested loops over genomes, contigs and surviving
 * pairs -- not the GPU design (2-bit words, canonical keys, counting sort).
 *
 * Parity pinning: the reference's own Python (unmodified bin/removeOutgroupPrimers.py) was run in the build
 * container against the pyahocorasick / Bio stand-ins (oracle/standins, oracle/run_reference.py) and this
 * restatement reproduces its surviving pairs, its (contig, size) sets and __extractKmerData's lists on the
 * committed fixtures (tests/golden/og_*.npz, made by oracle/make_golden_outgroup.py).  pyahocorasick's
 * iter() semantics (every occurrence of every word, by end index) are RECALLED, not checked against a real install.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint64_t n_hits, n_products;
    uint32_t *hits;        /* 5 per hit: key, genome, contig, start, strand; per (genome, contig): '+' then '-', scan order */
    uint32_t *products;    /* 4 per product: pair, genome, contig, size; distinct; pairs that are kept only */
    uint32_t *removed_at;  /* per pair: the genome pass that deleted it (:281-287), 0xFFFFFFFF = kept */
    char err[256];
} pfo_og_result;

typedef struct { uint32_t key, start; } og_hit;
typedef struct { og_hit *v; size_t n, cap; } hitvec;

static void hv_push(hitvec *h, uint32_t key, uint32_t start) {
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 64; h->v = (og_hit *)realloc(h->v, h->cap * sizeof(og_hit)); }
    h->v[h->n].key = key; h->v[h->n].start = start; h->n++;
}

static uint64_t fnv(const uint8_t *s, int n) {
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < n; i++) { h ^= s[i]; h *= 1099511628211ULL; }
    return h;
}

/* key strings of one length in a chained hash map */
typedef struct { int k; uint32_t nb; int32_t *head, *next; } keymap;

static uint8_t comp_byte(uint8_t c) {     /* Bio.Seq complement on the bytes that can match a primer */
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}

/* kmers.iter(seq) (:49): every occurrence of every key, by increasing end index (longest first at one end) */
static void scan_string(const uint8_t *s, size_t n, int strand, const keymap *maps, int nmaps, const uint8_t *const *keys,
                        hitvec *out) {
    for (size_t end = 0; end < n; end++) {
        for (int m = nmaps - 1; m >= 0; m--) {           /* maps are sorted by length */
            const int k = maps[m].k;
            if (end + 1 < (size_t)k) continue;
            const uint8_t *w = s + end + 1 - k;
            const uint64_t h = fnv(w, k);
            for (int32_t j = maps[m].head[h % maps[m].nb]; j >= 0; j = maps[m].next[j]) {
                if (memcmp(keys[j], w, (size_t)k) == 0) {
                    size_t start = end + 1 - (size_t)k;                    /* :51 */
                    if (strand) start = n - start - 1;                     /* :54-55 */
                    hv_push(out, (uint32_t)j, (uint32_t)start);
                }
            }
        }
    }
}

static int cmp_hit(const void *a, const void *b) {
    const og_hit *x = (const og_hit *)a, *y = (const og_hit *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->start < y->start ? -1 : x->start > y->start;
}
static int cmp_prod(const void *a, const void *b) {
    const uint32_t *x = (const uint32_t *)a, *y = (const uint32_t *)b;
    for (int i = 0; i < 4; i++)
        if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}

/* the starts of `key` in a list sorted by (key, start): [lo, hi) */
static void key_range(const og_hit *v, size_t n, uint32_t key, size_t *lo, size_t *hi) {
    size_t a = 0, b = n;
    while (a < b) { size_t m = (a + b) / 2; if (v[m].key < key) a = m + 1; else b = m; }
    *lo = a; b = n;
    while (a < b) { size_t m = (a + b) / 2; if (v[m].key <= key) a = m + 1; else b = m; }
    *hi = a;
}

void pfo_og_free(pfo_og_result *r) {
    if (!r) return;
    free(r->hits); free(r->products); free(r->removed_at); free(r);
}

/* keys: n_keys strings (key_len[i] bytes at key_bytes + key_off[i]); pairs: indices of str(fwd), str(rev);
 * genomes: like pfo_run (contigs concatenated, offsets per genome) */
pfo_og_result *pfo_outgroup(uint32_t n_keys, const uint8_t *key_bytes, const uint64_t *key_off, const uint8_t *key_len,
                            uint32_t n_pairs, const uint32_t *pair_fwd, const uint32_t *pair_rev, int32_t bad_lo, int32_t bad_hi,
                            int32_t n_genomes, const uint8_t *const *bases, const int64_t *const *contig_off,
                            const int32_t *n_contigs) {
    pfo_og_result *res = (pfo_og_result *)calloc(1, sizeof(pfo_og_result));
    /* __getAutomaton (:14-31): the distinct strings, here one chained map per length */
    const uint8_t **keys = (const uint8_t **)malloc(sizeof(uint8_t *) * (n_keys + 1));
    int present[256]; memset(present, 0, sizeof(present));
    for (uint32_t i = 0; i < n_keys; i++) { keys[i] = key_bytes + key_off[i]; present[key_len[i]]++; }
    keymap maps[256]; int nmaps = 0;
    for (int k = 1; k < 256; k++) {
        if (!present[k]) continue;
        keymap *m = &maps[nmaps++];
        m->k = k; m->nb = (uint32_t)(2 * present[k] + 1);
        m->head = (int32_t *)malloc(sizeof(int32_t) * m->nb);
        m->next = (int32_t *)malloc(sizeof(int32_t) * (n_keys + 1));
        for (uint32_t b = 0; b < m->nb; b++) m->head[b] = -1;
        for (uint32_t i = 0; i < n_keys; i++) {
            if (key_len[i] != k) continue;
            const uint64_t h = fnv(keys[i], k) % m->nb;
            int dup = 0;
            for (int32_t j = m->head[h]; j >= 0; j = m->next[j]) if (memcmp(keys[j], keys[i], (size_t)k) == 0) dup = 1;
            if (dup) { snprintf(res->err, sizeof(res->err), "duplicate key string"); return res; }
            m->next[i] = m->head[h]; m->head[h] = (int32_t)i;
        }
    }
    /* every contig of every genome, '+' and '-' (:209-226) */
    size_t total_contigs = 0;
    size_t *first = (size_t *)malloc(sizeof(size_t) * (size_t)(n_genomes + 1));
    for (int g = 0; g < n_genomes; g++) { first[g] = total_contigs; total_contigs += (size_t)n_contigs[g]; }
    first[n_genomes] = total_contigs;
    hitvec *plus = (hitvec *)calloc(total_contigs + 1, sizeof(hitvec)), *minus = (hitvec *)calloc(total_contigs + 1, sizeof(hitvec));
#pragma omp parallel for schedule(dynamic, 1)
    for (long t = 0; t < (long)total_contigs; t++) {
        int g = 0;
        while (first[g + 1] <= (size_t)t) g++;
        const size_t c = (size_t)t - first[g];
        const uint8_t *s = bases[g] + contig_off[g][c];
        const size_t n = (size_t)(contig_off[g][c + 1] - contig_off[g][c]);
        uint8_t *rc = (uint8_t *)malloc(n + 1);
        for (size_t i = 0; i < n; i++) rc[i] = comp_byte(s[n - 1 - i]);         /* contig.seq.reverse_complement() (:222) */
        scan_string(s, n, 0, maps, nmaps, keys, &plus[t]);
        scan_string(rc, n, 1, maps, nmaps, keys, &minus[t]);
        free(rc);
    }
    /* the hit lists in scan order */
    for (size_t t = 0; t < total_contigs; t++) res->n_hits += plus[t].n + minus[t].n;
    res->hits = (uint32_t *)malloc(sizeof(uint32_t) * 5 * (res->n_hits + 1));
    {
        size_t at = 0;
        for (int g = 0; g < n_genomes; g++)
            for (size_t c = 0; c < (size_t)n_contigs[g]; c++)
                for (int strand = 0; strand < 2; strand++) {
                    const hitvec *h = strand ? &minus[first[g] + c] : &plus[first[g] + c];
                    for (size_t i = 0; i < h->n; i++) {
                        uint32_t *o = res->hits + 5 * at++;
                        o[0] = h->v[i].key; o[1] = (uint32_t)g; o[2] = (uint32_t)c; o[3] = h->v[i].start; o[4] = (uint32_t)strand;
                    }
                }
    }
    /* the dict look-ups of :107-117 become binary searches in lists sorted by key */
    for (size_t t = 0; t < total_contigs; t++) {
        qsort(plus[t].v, plus[t].n, sizeof(og_hit), cmp_hit);
        qsort(minus[t].v, minus[t].n, sizeof(og_hit), cmp_hit);
    }
    /* the filtering loop (:236-291): genomes, contigs, the pairs still present */
    res->removed_at = (uint32_t *)malloc(sizeof(uint32_t) * (n_pairs + 1));
    for (uint32_t p = 0; p < n_pairs; p++) res->removed_at[p] = 0xFFFFFFFFu;
    size_t pcap = 1024, pn = 0;
    uint32_t *prod = (uint32_t *)malloc(sizeof(uint32_t) * 4 * pcap);
    for (int g = 0; g < n_genomes; g++)
        for (size_t c = 0; c < (size_t)n_contigs[g]; c++) {
            const hitvec *P = &plus[first[g] + c], *M = &minus[first[g] + c];
            if (P->n == 0 || M->n == 0) continue;
            for (uint32_t p = 0; p < n_pairs; p++) {
                if (res->removed_at[p] != 0xFFFFFFFFu) continue;             /* deleted from the dict earlier */
                const size_t pn0 = pn;
                int done = 0;
                for (int orient = 0; orient < 2 && !done; orient++) {      /* (:107-108) then (:116-117) */
                    const uint32_t a = orient ? pair_rev[p] : pair_fwd[p], b = orient ? pair_fwd[p] : pair_rev[p];
                    size_t a0, a1, b0, b1;
                    key_range(P->v, P->n, a, &a0, &a1);
                    key_range(M->v, M->n, b, &b0, &b1);
                    for (size_t x = a0; x < a1 && !done; x++)
                        for (size_t y = b0; y < b1; y++) {
                            const int64_t len = (int64_t)M->v[y].start - (int64_t)P->v[x].start + 1;   /* :80 */
                            if (len <= 0) continue;                                                     /* :83 */
                            if (len >= bad_lo && len <= bad_hi) { done = 1; break; }                    /* :283 */
                            if (pn == pcap) { pcap *= 2; prod = (uint32_t *)realloc(prod, sizeof(uint32_t) * 4 * pcap); }
                            prod[4 * pn] = p; prod[4 * pn + 1] = (uint32_t)g; prod[4 * pn + 2] = (uint32_t)c; prod[4 * pn + 3] = (uint32_t)len;
                            pn++;
                        }
                }
                if (done) { res->removed_at[p] = (uint32_t)g; pn = pn0; }   /* del pairs[(fwd, rev)] (:284); nothing saved for this contig */
            }
        }
    /* kept pairs only, as a set */
    size_t kept = 0;
    for (size_t i = 0; i < pn; i++)
        if (res->removed_at[prod[4 * i]] == 0xFFFFFFFFu) { memmove(prod + 4 * kept, prod + 4 * i, 16); kept++; }
    qsort(prod, kept, 16, cmp_prod);
    size_t uniq = 0;
    for (size_t i = 0; i < kept; i++)
        if (uniq == 0 || memcmp(prod + 4 * (uniq - 1), prod + 4 * i, 16) != 0) { memmove(prod + 4 * uniq, prod + 4 * i, 16); uniq++; }
    res->products = prod; res->n_products = uniq;
    for (size_t t = 0; t < total_contigs; t++) { free(plus[t].v); free(minus[t].v); }
    free(plus); free(minus); free(first);
    for (int m = 0; m < nmaps; m++) { free(maps[m].head); free(maps[m].next); }
    free(keys);
    return res;
}
