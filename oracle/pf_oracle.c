/*
 * pf_oracle.c -- CPU restatement of primerForge's candidate-k-mer stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under primerforge_b200/ may import, link
 * or execute this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs.  It deliberately follows the REFERENCE's algorithm (string-set algebra
 * over two count-min sketches per genome and k), not the GPU design, so that
 * the two are independent.
 *
 * Parity pinning: the reference's own Python (bin/getCandidateKmers.py,
 * bin/Primer.py) was run in the build container against stand-ins for its four
 * absent third-party dependencies (oracle/standins, oracle/run_reference.py)
 * and this restatement reproduces its shared-k-mer dict (keys, insertion
 * order, coordinates, strands) and candidate sets on the committed fixtures
 * (tests/golden, made by oracle/make_golden.py).  Tm is pinned by the six
 * README rows (README.md:375-377).  The internals of khmer / pyahocorasick are
 * RECALLED, not checked against a real install: parity for those two
 * semantics is "unpinned" (see DESIGN.md).
 *
 * Citations are into /root/reference/.
 *
 * Contract: bases are bytes; exactly 'A','C','G','T' (upper case) are
 * nucleotides, every other byte hashes like khmer's "everything else" (code
 * 3) and makes the windows that contain it ineligible (the reference would keep
 * such strings in its sets but they can never become candidates because
 * primer3.calc_tm rejects them; restricting the sets to ACGT-only strings
 * leaves the candidate set unchanged).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t k_min, k_max;          /* params.minLen / maxLen (getCandidateKmers.py:103) */
    uint64_t table_size;           /* khmer prime; Countgraph(k, 1e7, 1) -> 9999991 (:46-47) */
    double min_gc, max_gc;         /* :303 */
    double min_tm, max_tm;         /* :307 */
    int32_t max_repeat;            /* homopolymer length that is disallowed (:406-409) */
    int32_t thal_mode;             /* 0 = hairpin/homodimer always pass, 1 = pseudo (oracle/standins/primer3) */
    double temp_tolerance;         /* :342 */
    double mv, dv, dntp, dna, dmso, dmso_fact, formamide; /* primer3.calc_tm defaults (Primer.py:104) */
    int32_t n_threads;
    int32_t _pad;
} pfo_params;

typedef struct {
    uint64_t n_shared;
    int32_t n_genomes;
    /* per shared k-mer, in the reference's dict insertion order */
    uint8_t *k;          /* length */
    uint64_t *word;      /* 2-bit forward word of the key string (A0 T1 C2 G3, first base most significant) */
    uint8_t *flags;      /* bit0 gc ok, bit1 tm ok, bit2 no long repeat, bit3 picked (in the final candidate set) */
    int32_t *gc_count;
    double *tm;
    /* per genome x shared: contig index, leftmost (+) start, strand (0 '+', 1 '-') */
    int32_t *contig;
    int64_t *start;
    uint8_t *strand;
    uint64_t *n_allowed; /* [n_genomes * nk] size of each per-(genome,k) allowed set (after F-R / F^R) */
    char err[256];
} pfo_result;

/* ---------------------------------------------------------------- hashing */
/* khmer twobit_repr / twobit_comp (recalled, see header) */
static inline uint64_t code_fwd(uint8_t c) { return c == 'A' ? 0 : c == 'T' ? 1 : c == 'C' ? 2 : 3; }
static inline uint64_t code_cmp(uint8_t c) { return c == 'A' ? 1 : c == 'T' ? 0 : c == 'C' ? 3 : 2; }
static inline int is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static inline uint8_t comp_base(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return c; }
}

/* ---------------------------------------------------------------- small utilities */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

typedef struct { uint64_t *v; size_t n, cap; } vec64;
static void v_push(vec64 *v, uint64_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->v = (uint64_t *)realloc(v->v, v->cap * 8); }
    v->v[v->n++] = x;
}

/* open addressing map: 64-bit key -> 32-bit index */
typedef struct { uint64_t *keys; uint32_t *vals; uint64_t mask; } map64;
static inline uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
#define MAP_EMPTY 0xFFFFFFFFFFFFFFFFULL
static void map_init(map64 *m, size_t n) {
    uint64_t cap = 16; while (cap < n * 2 + 2) cap <<= 1;
    m->keys = (uint64_t *)malloc(cap * 8); m->vals = (uint32_t *)malloc(cap * 4); m->mask = cap - 1;
    memset(m->keys, 0xFF, cap * 8);
}
static void map_put(map64 *m, uint64_t key, uint32_t val) {
    uint64_t i = mix64(key) & m->mask;
    while (m->keys[i] != MAP_EMPTY && m->keys[i] != key) i = (i + 1) & m->mask;
    m->keys[i] = key; m->vals[i] = val;
}
static inline int64_t map_get(const map64 *m, uint64_t key) {
    uint64_t i = mix64(key) & m->mask;
    while (m->keys[i] != MAP_EMPTY) { if (m->keys[i] == key) return m->vals[i]; i = (i + 1) & m->mask; }
    return -1;
}
static void map_free(map64 *m) { free(m->keys); free(m->vals); m->keys = NULL; m->vals = NULL; }

/* ---------------------------------------------------------------- a-1: __getAllowedKmers (getCandidateKmers.py:15-65) */
/* consume(): every window of the joined string, no validation (:50-51) */
static void sketch_consume(const uint8_t *s, size_t n, int k, uint64_t P, uint8_t *tab) {
    if (n < (size_t)k) return;
    const uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t h = 0, r = 0;
    for (size_t i = 0; i < n; i++) {
        h = ((h << 2) | code_fwd(s[i])) & mask;
        r = (r >> 2) | (code_cmp(s[i]) << (2 * (k - 1)));
        if (i + 1 >= (size_t)k) {
            uint64_t c = h < r ? h : r;           /* uniqify_rc */
            uint8_t *b = &tab[c % P];
            if (*b != 255) (*b)++;                /* one-byte saturating counter */
        }
    }
}

/* the set comprehension (:54-55): isOneEndGc and get()==1 and no junction char;
 * additionally ACGT-only (see header).  Emits the forward 2-bit word of each string. */
static void sketch_select(const uint8_t *s, size_t n, int k, uint64_t P, const uint8_t *tab, vec64 *out) {
    if (n < (size_t)k) return;
    const uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t h = 0, r = 0;
    size_t run = 0; /* consecutive ACGT bytes ending here */
    for (size_t i = 0; i < n; i++) {
        h = ((h << 2) | code_fwd(s[i])) & mask;
        r = (r >> 2) | (code_cmp(s[i]) << (2 * (k - 1)));
        run = is_acgt(s[i]) ? run + 1 : 0;
        if (i + 1 >= (size_t)k && run >= (size_t)k) {
            uint8_t first = s[i + 1 - k], last = s[i];
            int one_end_gc = first == 'G' || first == 'C' || last == 'G' || last == 'C';   /* :29-33 */
            if (!one_end_gc) continue;
            uint64_t c = h < r ? h : r;
            if (tab[c % P] == 1) v_push(out, h);  /* :35-38 */
        }
    }
}

/* sorted-set algebra for :58-63 */
static size_t set_difference(const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *o) {
    size_t i = 0, j = 0, n = 0;
    while (i < na) {
        while (j < nb && b[j] < a[i]) j++;
        if (j < nb && b[j] == a[i]) { i++; continue; }
        o[n++] = a[i++];
    }
    return n;
}
static size_t set_symdiff(const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *o) {
    size_t i = 0, j = 0, n = 0;
    while (i < na || j < nb) {
        if (j >= nb || (i < na && a[i] < b[j])) o[n++] = a[i++];
        else if (i >= na || b[j] < a[i]) o[n++] = b[j++];
        else { i++; j++; }
    }
    return n;
}
static size_t set_intersect(uint64_t *a, size_t na, const uint64_t *b, size_t nb) {
    size_t i = 0, j = 0, n = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (b[j] < a[i]) j++;
        else { a[n++] = a[i]; i++; j++; }
    }
    return n;
}
static size_t sort_unique(uint64_t *v, size_t n) {
    if (!n) return 0;
    qsort(v, n, 8, cmp_u64);
    size_t m = 1;
    for (size_t i = 1; i < n; i++) if (v[i] != v[m - 1]) v[m++] = v[i];
    return m;
}

/* ---------------------------------------------------------------- a-6: Primer GC / Tm (Primer.py:89-104) */
static const int NN_H[16] = { /* index = 4*code(first)+code(second), codes A0 T1 C2 G3 */
    /*AA*/79, /*AT*/72, /*AC*/84, /*AG*/78,
    /*TA*/72, /*TT*/79, /*TC*/82, /*TG*/85,
    /*CA*/85, /*CT*/78, /*CC*/80, /*CG*/106,
    /*GA*/82, /*GT*/84, /*GC*/98, /*GG*/80 };
static const int NN_S[16] = {
    /*AA*/222, /*AT*/204, /*AC*/224, /*AG*/210,
    /*TA*/213, /*TT*/222, /*TC*/222, /*TG*/227,
    /*CA*/227, /*CT*/210, /*CC*/199, /*CG*/272,
    /*GA*/222, /*GT*/224, /*GC*/244, /*GG*/199 };

static inline int base_at(uint64_t w, int k, int i) { return (int)((w >> (2 * (k - 1 - i))) & 3); }

static uint64_t revcomp_word(uint64_t w, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | ((w & 3) ^ 1); w >>= 2; }
    return r;
}

/* primer3 oligotm(), SantaLucia table + SantaLucia salt correction, as restated in
 * oracle/standins/primer3 (operation order kept identical so doubles match bit for bit) */
static double oligo_tm(uint64_t w, int k, const pfo_params *p, int *gc_out) {
    int dh = 0, ds = 0, gc = 0;
    int sym = (k % 2 == 0) && revcomp_word(w, k) == w;
    if (sym) ds += 14;
    int ends[2] = { base_at(w, k, 0), base_at(w, k, k - 1) };
    for (int e = 0; e < 2; e++) {
        if (ends[e] < 2) { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    }
    for (int i = 0; i < k - 1; i++) {
        int idx = 4 * base_at(w, k, i) + base_at(w, k, i + 1);
        dh += NN_H[idx]; ds += NN_S[idx];
    }
    for (int i = 0; i < k; i++) gc += base_at(w, k, i) >= 2;
    double delta_h = dh * -100.0;
    double delta_s = ds * -0.1;
    double dv = p->dv, dntp = p->dntp;
    if (dv == 0) dntp = 0;
    if (dv < dntp) dv = dntp;
    double k_mm = p->mv + 120.0 * sqrt(dv - dntp);
    delta_s = delta_s + 0.368 * (k - 1) * log(k_mm / 1000.0);
    double tm;
    if (sym) tm = delta_h / (delta_s + 1.987 * log(p->dna / 1000000000.0)) - 273.15;
    else     tm = delta_h / (delta_s + 1.987 * log(p->dna / 4000000000.0)) - 273.15;
    tm -= p->dmso * p->dmso_fact;
    tm += (0.453 * gc / k - 2.88) * p->formamide;   /* oligotm.c: 0.453 * GC_count / len, left to right */
    if (gc_out) *gc_out = gc;
    return tm;
}

static int has_long_repeat(uint64_t w, int k, int max_repeat) { /* :309-320, :406-409 */
    int run = 1;
    if (max_repeat <= 1) return 1;
    for (int i = 1; i < k; i++) {
        run = base_at(w, k, i) == base_at(w, k, i - 1) ? run + 1 : 1;
        if (run >= max_repeat) return 1;
    }
    return 0;
}

/* pseudo thal (oracle/standins/primer3/__init__.py:_pseudo_thal) */
static uint32_t crc32_tab[256];
static void crc32_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int j = 0; j < 8; j++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc32_tab[i] = c;
    }
}
static double pseudo_thal(uint64_t w, int k, uint32_t salt) {
    static const char L[4] = { 'A', 'T', 'C', 'G' };
    uint32_t c = 0xFFFFFFFFu;
    for (int i = 0; i < k; i++) c = crc32_tab[(c ^ (uint8_t)L[base_at(w, k, i)]) & 0xFF] ^ (c >> 8);
    c ^= 0xFFFFFFFFu;
    return (double)((c ^ (salt * 0x9E3779B1u)) % 6400u) / 100.0;
}

/* ---------------------------------------------------------------- driver */
typedef struct { int32_t contig; int64_t start; uint8_t strand; uint8_t set; } entry_t;

void pfo_free(pfo_result *r) {
    if (!r) return;
    free(r->k); free(r->word); free(r->flags); free(r->gc_count); free(r->tm);
    free(r->contig); free(r->start); free(r->strand); free(r->n_allowed);
    free(r);
}

/*
 * bases[g]        : concatenated contigs of genome g (no separators)
 * contig_off[g]   : n_contigs[g]+1 offsets into bases[g]
 * genome 0 is "the first genome" (getCandidateKmers.py:85,108)
 */
pfo_result *pfo_run(int32_t n_genomes, const uint8_t *const *bases, const int64_t *const *contig_off,
                    const int32_t *n_contigs, const pfo_params *p) {
    pfo_result *res = (pfo_result *)calloc(1, sizeof(pfo_result));
    res->n_genomes = n_genomes;
    const int kmin = p->k_min, kmax = p->k_max, nk = kmax - kmin + 1;
    const uint64_t P = p->table_size;
    if (kmin < 1 || kmax > 32 || nk < 1 || n_genomes < 1) { snprintf(res->err, sizeof res->err, "bad parameters"); return res; }
    crc32_init();
#ifdef _OPENMP
    if (p->n_threads > 0) omp_set_num_threads(p->n_threads);
#endif
    res->n_allowed = (uint64_t *)calloc((size_t)n_genomes * nk, 8);

    /* ---- a-2 generateArgs (:81-108): joined strings, junction = '~' * maxLen */
    uint8_t **fwd = (uint8_t **)calloc(n_genomes, sizeof(void *));
    uint8_t **rev = (uint8_t **)calloc(n_genomes, sizeof(void *));
    size_t *jlen = (size_t *)calloc(n_genomes, sizeof(size_t));
    for (int g = 0; g < n_genomes; g++) {
        int C = n_contigs[g];
        size_t tot = (size_t)contig_off[g][C] + (size_t)(C > 0 ? C - 1 : 0) * kmax;
        jlen[g] = tot;
        fwd[g] = (uint8_t *)malloc(tot + 1); rev[g] = (uint8_t *)malloc(tot + 1);
        size_t o = 0;
        for (int c = 0; c < C; c++) {
            int64_t a = contig_off[g][c], b = contig_off[g][c + 1];
            if (c) { memset(fwd[g] + o, '~', kmax); memset(rev[g] + o, '~', kmax); o += kmax; }
            for (int64_t i = a; i < b; i++) {
                fwd[g][o + (i - a)] = bases[g][i];
                rev[g][o + (i - a)] = comp_base(bases[g][b - 1 - (i - a)]);  /* rec.seq.reverse_complement() (:95) */
            }
            o += (size_t)(b - a);
        }
    }

    /* ---- a-1 for every (genome, k) task (:115-118) */
    vec64 *allowed = (vec64 *)calloc((size_t)n_genomes * nk, sizeof(vec64));
    int ntask = n_genomes * nk;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < ntask; t++) {
        int g = t / nk, k = kmin + t % nk;
        uint8_t *tf = (uint8_t *)calloc(P, 1), *tr = (uint8_t *)calloc(P, 1);
        sketch_consume(fwd[g], jlen[g], k, P, tf);   /* fkh.consume(fwd) (:50) */
        sketch_consume(rev[g], jlen[g], k, P, tr);   /* rkh.consume(rev) (:51) */
        vec64 F = {0}, R = {0};
        sketch_select(fwd[g], jlen[g], k, P, tf, &F);   /* :54 */
        sketch_select(rev[g], jlen[g], k, P, tr, &R);   /* :55 */
        free(tf); free(tr);
        F.n = sort_unique(F.v, F.n); R.n = sort_unique(R.v, R.n);
        vec64 out = {0};
        out.v = (uint64_t *)malloc((F.n + R.n + 1) * 8);
        if (g == 0) out.n = set_difference(F.v, F.n, R.v, R.n, out.v);       /* :58-59 */
        else        out.n = set_symdiff(F.v, F.n, R.v, R.n, out.v);          /* :62-63 */
        free(F.v); free(R.v);
        allowed[t] = out;
        res->n_allowed[t] = out.n;
    }

    /* ---- per-k intersection over genomes (:127-142) */
    vec64 *shared_k = (vec64 *)calloc(nk, sizeof(vec64));
    size_t n_total = 0;
    for (int ki = 0; ki < nk; ki++) {
        vec64 s = allowed[ki];  /* genome 0 */
        for (int g = 1; g < n_genomes; g++) s.n = set_intersect(s.v, s.n, allowed[g * nk + ki].v, allowed[g * nk + ki].n);
        shared_k[ki] = s;
        n_total += s.n;
    }
    for (int g = 1; g < n_genomes; g++) for (int ki = 0; ki < nk; ki++) free(allowed[g * nk + ki].v);
    free(allowed);

    /* ---- a-3 locate (:160-219): automaton over all k; '+' scan then '-' scan of every contig of
     * every genome; at one end index the longest word is reported first; later match overwrites */
    map64 *maps = (map64 *)calloc(nk, sizeof(map64));
    size_t *base_idx = (size_t *)calloc(nk + 1, sizeof(size_t));
    for (int ki = 0; ki < nk; ki++) {
        base_idx[ki + 1] = base_idx[ki] + shared_k[ki].n;
        map_init(&maps[ki], shared_k[ki].n);
        for (size_t i = 0; i < shared_k[ki].n; i++) map_put(&maps[ki], shared_k[ki].v[i], (uint32_t)i);
    }
    entry_t *entries = (entry_t *)calloc((size_t)n_genomes * (n_total ? n_total : 1), sizeof(entry_t));
    int64_t *rank_of = (int64_t *)malloc((n_total ? n_total : 1) * 8);   /* insertion order */
    for (size_t i = 0; i < n_total; i++) rank_of[i] = -1;
    size_t *order = (size_t *)malloc((n_total ? n_total : 1) * sizeof(size_t));
    size_t n_ins = 0;

    /* genome 0 first and alone: it assigns the insertion ranks (every shared k-mer occurs on its (+)
     * strand); the other genomes only fill their own rows and run in parallel */
#pragma omp parallel for schedule(dynamic, 1)
    for (int g = 0; g < n_genomes; g++) {
        if (!n_total) continue;
        int C = n_contigs[g];
        for (int c = 0; c < C; c++) {
            int64_t a = contig_off[g][c], b = contig_off[g][c + 1], len = b - a;
            for (int strand = 0; strand < 2; strand++) {
                uint64_t h = 0;  /* rolling forward word of the last 32 bases of the scanned string */
                int64_t run = 0;
                for (int64_t e = 0; e < len; e++) {
                    uint8_t ch = strand == 0 ? bases[g][a + e] : comp_base(bases[g][b - 1 - e]);
                    h = (h << 2) | code_fwd(ch);
                    run = is_acgt(ch) ? run + 1 : 0;
                    for (int k = kmax; k >= kmin; k--) {          /* longest first at one end index */
                        if (run < k) continue;
                        uint64_t w = k == 32 ? h : (h & ((1ULL << (2 * k)) - 1));
                        int ki = k - kmin;
                        int64_t id = map_get(&maps[ki], w);
                        if (id < 0) continue;
                        size_t gi = base_idx[ki] + (size_t)id;
                        if (g == 0 && rank_of[gi] < 0) { rank_of[gi] = (int64_t)n_ins; order[n_ins++] = gi; }
                        entry_t *en = &entries[(size_t)g * n_total + gi];
                        en->contig = c;
                        en->start = strand == 0 ? e - k + 1 : len - e - 1;   /* :173-178 */
                        en->strand = (uint8_t)strand;
                        en->set = 1;
                    }
                }
            }
        }
    }

    /* ---- assemble result in insertion order */
    size_t N = n_ins;
    res->n_shared = N;
    size_t Na = N ? N : 1;
    res->k = (uint8_t *)calloc(Na, 1); res->word = (uint64_t *)calloc(Na, 8); res->flags = (uint8_t *)calloc(Na, 1);
    res->gc_count = (int32_t *)calloc(Na, 4); res->tm = (double *)calloc(Na, 8);
    res->contig = (int32_t *)calloc(Na * n_genomes, 4); res->start = (int64_t *)calloc(Na * n_genomes, 8);
    res->strand = (uint8_t *)calloc(Na * n_genomes, 1);
    int missing = 0;
#pragma omp parallel for schedule(static) reduction(+:missing)
    for (size_t r = 0; r < N; r++) {
        size_t gi = order[r];
        int ki = 0; while (gi >= base_idx[ki + 1]) ki++;
        int k = kmin + ki;
        uint64_t w = shared_k[ki].v[gi - base_idx[ki]];
        res->k[r] = (uint8_t)k; res->word[r] = w;
        int gc; double tm = oligo_tm(w, k, p, &gc);
        res->gc_count[r] = gc; res->tm[r] = tm;
        double gc_per = (double)gc / k * 100;                          /* Primer.py:100 */
        uint8_t f = 0;
        if (gc_per >= p->min_gc && gc_per <= p->max_gc) f |= 1;        /* :303 */
        if (tm >= p->min_tm && tm <= p->max_tm) f |= 2;                /* :307 */
        if (!has_long_repeat(w, k, p->max_repeat)) f |= 4;             /* :309-320 */
        res->flags[r] = f;
        for (int g = 0; g < n_genomes; g++) {
            entry_t *en = &entries[(size_t)g * n_total + gi];
            if (!en->set) missing++;
            res->contig[(size_t)g * N + r] = en->contig;
            res->start[(size_t)g * N + r] = en->start;
            res->strand[(size_t)g * N + r] = en->strand;
        }
    }
    if (missing) snprintf(res->err, sizeof res->err, "internal: %d shared k-mers were not located in every genome", missing);

    /* ---- a-4/a-5 per genome: group by (contig, start) in insertion order, first member that passes
     * GC, Tm, repeats, hairpin, homodimer is the pick (:222-247, :276-388).  The `seen` filter
     * (:250-273) only skips groups that would yield the same pick and is therefore omitted. */
    if (N) {
        uint8_t *pass = (uint8_t *)malloc(N);
        double thr = p->min_tm - p->temp_tolerance;                    /* :342-343, :367-368 */
#pragma omp parallel for schedule(static)
        for (size_t r = 0; r < N; r++) {
            int ok = (res->flags[r] & 7) == 7;
            if (ok && p->thal_mode == 1) {
                uint64_t w = res->word[r]; int k = res->k[r]; uint64_t rc = revcomp_word(w, k);
                ok = pseudo_thal(w, k, 1) < thr && pseudo_thal(rc, k, 1) < thr &&    /* noHairpins */
                     pseudo_thal(w, k, 2) < thr && pseudo_thal(rc, k, 2) < thr;      /* noHomodimers */
            }
            pass[r] = (uint8_t)ok;
        }
        for (int g = 0; g < n_genomes; g++) {
            /* first passing rank per (contig, start): a map keyed by contig * 2^40 + start */
            map64 grp; map_init(&grp, N);
            for (size_t r = 0; r < N; r++) {
                if (!pass[r]) continue;
                uint64_t key = ((uint64_t)res->contig[(size_t)g * N + r] << 40) | (uint64_t)res->start[(size_t)g * N + r];
                if (map_get(&grp, key) < 0) { map_put(&grp, key, (uint32_t)r); res->flags[r] |= 8; }
            }
            map_free(&grp);
        }
        free(pass);
    }

    for (int ki = 0; ki < nk; ki++) { map_free(&maps[ki]); free(shared_k[ki].v); }
    free(maps); free(shared_k); free(base_idx); free(entries); free(rank_of); free(order);
    for (int g = 0; g < n_genomes; g++) { free(fwd[g]); free(rev[g]); }
    free(fwd); free(rev); free(jlen);
    return res;
}

/* single-oligo helpers for the known-answer tests (README.md:375-377, primer_test.py) */
double pfo_calc_tm(const char *seq, const pfo_params *p) {
    int k = (int)strlen(seq); uint64_t w = 0;
    if (k < 2 || k > 32) return -999999.9999;
    for (int i = 0; i < k; i++) { if (!is_acgt((uint8_t)seq[i])) return -999999.9999; w = (w << 2) | code_fwd((uint8_t)seq[i]); }
    return oligo_tm(w, k, p, NULL);
}

int pfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
