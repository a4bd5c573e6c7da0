/*
 * pf_oracle_pairs.c -- CPU restatement of primerForge's pair search
 * (/root/reference/bin/getPrimerPairs.py::_getPrimerPairs, :490-564 and the functions it calls).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under primerforge_b200/ may import, link or execute this file; it is the
 * checker used by tests/ and by bench.py's cpu_baseline leg of the pair-search block.  It follows the REFERENCE's
 * control flow on lists -- the sequential binning loop, the per-bin length lists that __reduceBinSize edits while it
 * iterates over them, the nested bin-pair and primer-pair loops with their early exits -- not the GPU design.
 *
 * Parity pinning: the reference's own Python (unmodified bin/getPrimerPairs.py) was run in the build container
 * against the stand-in dependencies (oracle/standins: primer3.calc_heterodimer_tm returns 0.0, i.e. no pair is
 * rejected as a dimer) and this restatement reproduces its pairs (order, sequences) and every field of its Product
 * objects on the committed fixtures (tests/golden/pp_*.npz, made by oracle/make_golden_pairs.py).
 *
 * Input model: S candidate k-mers, the same set in every genome (getCandidateKmers.py:442-489): word = the k-mer
 * as it reads on the first genome's (+) strand (2 bits per base, A0 T1 C2 G3, first base most significant), length,
 * Tm.  Per genome the candidates in the order of its dict of lists (contigs in dict order, each list in list order):
 * candidate id, contig index (dense, in dict order), leftmost (+) coordinate, strand.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t max_bin;          /* params.maxBinSize */
    int32_t min_primer_len;   /* params.minLen */
    int32_t min_pcr, max_pcr; /* params.minPcr / maxPcr */
    double max_tm_diff;       /* params.maxTmDiff */
} pfo_pp_params;

typedef struct {
    uint64_t n_pairs;         /* pairs proposed by __getCandidatePrimerPairs' generator (:283-341), in its order */
    uint32_t *pairs;          /* 6 per pair: fwd candidate, rev candidate, contig, fbin, rbin, pcrLen (first genome) */
    uint32_t *shared;         /* [n_pairs][n_genomes][4]: valid, length, fbin, rbin (__getAllSharedPrimerPairs :367-464 +
                                 __updateBinsForUnprocessedGenomes :467-502); genome 0 = the first genome (always valid) */
    uint32_t *bin_of;         /* [n_genomes][S]: bin number of every candidate within its contig (__binCandidateKmers) */
    uint8_t *reduced;         /* [S]: 1 = removed from its bin by __reduceBinSize (first genome) */
    char err[256];
} pfo_pp_result;

static inline uint64_t rc_word(uint64_t w, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | ((w & 3u) ^ 1u); w >>= 2; }
    return r;
}

/* p1.seq in p2.seq (:34) on 2-bit words */
static int contains(uint64_t big, int kb, uint64_t small, int ks) {
    const uint64_t mask = ks == 32 ? ~0ull : ((1ull << (2 * ks)) - 1);
    for (int o = 0; o + ks <= kb; o++)
        if (((big >> (2 * (kb - ks - o))) & mask) == small) return 1;
    return 0;
}

typedef struct { uint32_t *v; int n; } ilist;

static void il_remove(ilist *l, uint32_t x) {      /* list.remove: the first equal element */
    for (int i = 0; i < l->n; i++)
        if (l->v[i] == x) { memmove(l->v + i, l->v + i + 1, sizeof(uint32_t) * (size_t)(l->n - i - 1)); l->n--; return; }
}

void pfo_pp_free(pfo_pp_result *r) {
    if (!r) return;
    free(r->pairs); free(r->shared); free(r->bin_of); free(r->reduced); free(r);
}

/* __binCandidateKmers (:39-98) for one genome: bin number of every candidate; '-' candidates are reverse
 * complemented first (:63-64), which leaves their leftmost coordinate and length, i.e. curStart / curEnd */
static void bin_genome(uint64_t S, const uint32_t *id, const uint32_t *contig, const uint32_t *left, const uint8_t *k,
                       int32_t max_bin, uint32_t *bin_of) {
    uint64_t p = 0;
    while (p < S) {
        const uint32_t c = contig[p];
        int64_t currentBin = 0, prevEnd = -1, binStart = 0;
        int first = 1;
        for (; p < S && contig[p] == c; p++) {
            const int64_t curStart = left[p], curEnd = (int64_t)left[p] + k[id[p]] - 1;
            const int64_t binLen = curEnd - binStart;                                   /* :71 */
            if (first) { prevEnd = curEnd; binStart = curStart; first = 0; }            /* :74-77 */
            else if (curStart < prevEnd && binLen <= max_bin) { prevEnd = curEnd; }     /* :80-82 */
            else { prevEnd = curEnd; binStart = curStart; currentBin++; }               /* :85-90 */
            bin_of[id[p]] = (uint32_t)currentBin;
        }
    }
}

/* genomes[g]: id / contig / left / strand arrays of length S in the genome's list order */
pfo_pp_result *pfo_pairs(uint64_t S, const uint64_t *word, const uint8_t *k, const double *tm, int32_t n_genomes,
                         const uint32_t *const *g_id, const uint32_t *const *g_contig, const uint32_t *const *g_left,
                         const uint8_t *const *g_strand, const pfo_pp_params *prm) {
    pfo_pp_result *res = (pfo_pp_result *)calloc(1, sizeof(pfo_pp_result));
    const int G = n_genomes;
    res->bin_of = (uint32_t *)calloc((size_t)G * S + 1, sizeof(uint32_t));
    res->reduced = (uint8_t *)calloc(S + 1, 1);
    for (int g = 0; g < G; g++) bin_genome(S, g_id[g], g_contig[g], g_left[g], k, prm->max_bin, res->bin_of + (size_t)g * S);

    /* ---- first genome: the bins as lists, __reduceBinSize (:10-36) */
    const uint32_t *id0 = g_id[0], *ct0 = g_contig[0], *lf0 = g_left[0];
    /* bins in (contig, bin number) order = list-position order; bin b covers positions [bs[b], bs[b + 1]) */
    uint64_t nb = 0;
    uint64_t *bs = (uint64_t *)malloc(sizeof(uint64_t) * (S + 2));
    for (uint64_t p = 0; p < S; p++)
        if (p == 0 || ct0[p] != ct0[p - 1] || res->bin_of[id0[p]] != res->bin_of[id0[p - 1]]) bs[nb++] = p;
    bs[nb] = S;
    /* the members of every bin after the reduction, in list order */
    uint32_t *mem = (uint32_t *)malloc(sizeof(uint32_t) * (S + 1));   /* candidate ids, bin b at mem[ms[b] .. ms[b] + mn[b]) */
    uint64_t *ms = (uint64_t *)malloc(sizeof(uint64_t) * (nb + 1));
    int *mn = (int *)malloc(sizeof(int) * (nb + 1));
    for (uint64_t b = 0; b < nb; b++) {
        const int n = (int)(bs[b + 1] - bs[b]);
        ilist bin; bin.v = mem + bs[b]; bin.n = n;
        for (int i = 0; i < n; i++) bin.v[i] = id0[bs[b] + i];
        /* lengths = dict len -> list of primers, in bin order (:20-23) */
        ilist lens[33];
        uint32_t *store = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
        int at = 0;
        for (int L = 0; L <= 32; L++) {
            lens[L].v = store + at; lens[L].n = 0;
            for (int i = 0; i < n; i++) if (k[bin.v[i]] == L) lens[L].v[lens[L].n++] = bin.v[i];
            at += lens[L].n;
        }
        /* for small, big in combinations(sorted(lengths.keys()), 2) (:26) */
        for (int small = 0; small <= 32; small++) {
            /* (a length that is not a key of the dict is an empty list here: its loops do nothing) */
            for (int big = small + 1; big <= 32; big++) {
                /* for p1 in lengths[small] (:28) -- a list iterator: after a removal at index i the next element read
                 * is the one that was at i + 2 */
                for (int i = 0; i < lens[small].n; i++) {
                    const uint32_t p1 = lens[small].v[i];
                    for (int j = 0; j < lens[big].n; j++) {
                        const uint32_t p2 = lens[big].v[j];
                        if (contains(word[p2], k[p2], word[p1], k[p1])) {        /* :34 */
                            il_remove(&bin, p1);                                   /* :35 */
                            il_remove(&lens[small], p1);                           /* :36 */
                            res->reduced[p1] = 1;
                            break;                                                 /* (the for over lengths[small] goes on at i + 1) */
                        }
                    }
                }
            }
        }
        free(store);
        ms[b] = bs[b]; mn[b] = bin.n;
    }

    /* ---- __getBinPairs (:101-143) + the generator of __getCandidatePrimerPairs (:283-341) */
    size_t cap = 1024;
    res->pairs = (uint32_t *)malloc(sizeof(uint32_t) * 6 * cap);
    /* inverse order of the first genome: candidate id -> leftmost */
    int64_t *left_of = (int64_t *)malloc(sizeof(int64_t) * (S + 1));
    for (uint64_t p = 0; p < S; p++) left_of[id0[p]] = lf0[p];
    for (uint64_t b1 = 0; b1 + 1 < nb; b1++) {
        const uint32_t c = ct0[bs[b1]];
        const uint32_t *B1 = mem + ms[b1]; const int n1 = mn[b1];
        for (uint64_t b2 = b1 + 1; b2 < nb && ct0[bs[b2]] == c; b2++) {
            const uint32_t *B2 = mem + ms[b2]; const int n2 = mn[b2];
            const int64_t smallest = (left_of[B2[0]] + prm->min_primer_len) -
                                     ((left_of[B1[n1 - 1]] + k[B1[n1 - 1]] - 1) - prm->min_primer_len);            /* :128 */
            const int64_t largest = (left_of[B2[n2 - 1]] + k[B2[n2 - 1]] - 1) - left_of[B1[0]];                 /* :131 */
            if (smallest > prm->max_pcr) break;                                                                 /* :134-135 */
            if (largest < prm->min_pcr) continue;                                                               /* :138-139 */
            /* the first suitable (fwd, rev) of the bin pair (:297-341) */
            int found = 0;
            for (int i = 0; i < n1 && !found; i++) {
                const uint32_t f = B1[i];
                if ((word[f] & 2u) == 0) continue;                       /* fwd.seq[-1] in {G, C} (:271-272): codes C2 G3 */
                for (int j = 0; j < n2; j++) {
                    const uint32_t r = B2[j];
                    if (((word[r] >> (2 * (k[r] - 1))) & 2u) == 0) continue;       /* rev.seq[0] in {G, C} (:273-274) */
                    const int64_t pcr = (left_of[r] + k[r] - 1) - left_of[f] + 1;                               /* :221 */
                    if (pcr < prm->min_pcr || pcr > prm->max_pcr) continue;                                     /* :224 */
                    if (!(fabs(tm[f] - tm[r]) <= prm->max_tm_diff)) continue;                                   /* :227 */
                    if (res->n_pairs == cap) { cap *= 2; res->pairs = (uint32_t *)realloc(res->pairs, sizeof(uint32_t) * 6 * cap); }
                    uint32_t *o = res->pairs + 6 * res->n_pairs++;
                    o[0] = f; o[1] = r; o[2] = c; o[3] = res->bin_of[f]; o[4] = res->bin_of[r]; o[5] = (uint32_t)pcr;
                    found = 1;
                    break;
                }
            }
        }
    }

    /* ---- __getAllSharedPrimerPairs (:367-464) + __updateBinsForUnprocessedGenomes (:467-502) */
    res->shared = (uint32_t *)calloc((size_t)res->n_pairs * (size_t)G * 4 + 4, sizeof(uint32_t));
    for (int g = 0; g < G; g++) {
        /* kmers[name][seq] -> Primer (:349-362): candidate id -> (contig, leftmost, strand) in this genome */
        uint32_t *ct = (uint32_t *)malloc(sizeof(uint32_t) * (S + 1));
        int64_t *lf = (int64_t *)malloc(sizeof(int64_t) * (S + 1));
        uint8_t *st = (uint8_t *)malloc(S + 1);
        for (uint64_t p = 0; p < S; p++) { const uint32_t s = g_id[g][p]; ct[s] = g_contig[g][p]; lf[s] = g_left[g][p]; st[s] = g_strand[g][p]; }
        const uint32_t *bin_g = res->bin_of + (size_t)g * S;
        for (uint64_t q = 0; q < res->n_pairs; q++) {
            const uint32_t *pr = res->pairs + 6 * q;
            uint32_t *o = res->shared + (q * (size_t)G + (size_t)g) * 4;
            const uint32_t a = pr[0], b = pr[1];
            if (g == 0) { o[0] = 1; o[1] = pr[5]; o[2] = pr[3]; o[3] = pr[4]; continue; }       /* out[(p1, p2)][firstName] = product (:398) */
            /* p1.seq is the first candidate's own string: found as it is (k1_rev False); p2.seq is the reverse
             * complement of the second candidate's string: found as it is only when that string is its own reverse
             * complement (:403-431) */
            const int k1_rev = 0;
            const int k2_rev = !(rc_word(word[b], k[b]) == word[b]);
            if (ct[a] != ct[b]) continue;                                                       /* :434 */
            if (k1_rev == k2_rev) continue;                                                     /* :436 */
            /* Primer.start: leftmost on (+), rightmost on (-) (Primer.py:42-45) */
            const int64_t s1 = st[a] ? lf[a] + k[a] - 1 : lf[a], s2 = st[b] ? lf[b] + k[b] - 1 : lf[b];
            const uint32_t fw = s1 < s2 ? a : b, rv = s1 < s2 ? b : a;                          /* :438-445 */
            if (st[fw] != st[rv]) continue;                                                     /* :448 */
            /* both on (+) after an optional reverseComplement of both (:450-452): start = leftmost, end = leftmost + len - 1 */
            const int64_t length = (lf[rv] + k[rv] - 1) - lf[fw] + 1;                           /* :455 */
            if (length < prm->min_pcr || length > prm->max_pcr) continue;                       /* :456 */
            o[0] = 1; o[1] = (uint32_t)length; o[2] = bin_g[a]; o[3] = bin_g[b];                /* :463, :497-502 */
        }
        free(ct); free(lf); free(st);
    }
    free(bs); free(mem); free(ms); free(mn); free(left_of);
    return res;
}
