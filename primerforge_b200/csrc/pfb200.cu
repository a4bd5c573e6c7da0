// pfb200.cu -- B200 (sm_100a) candidate-k-mer stage of primerForge, from scratch.
//
// What the reference computes (bin/getCandidateKmers.py:15-219, file:line citations are into
// /root/reference): per genome and per k two khmer count-min sketches (P = 9 999 991 one-byte bins,
// bin = min(fwd, rc) 2-bit word mod P) over the '~'-joined forward and reverse-complement strings,
// the k-mers whose bin reads exactly 1, set algebra between strands, a per-k intersection over
// genomes, an Aho-Corasick relocation of the survivors, grouping by position and GC / Tm / repeat
// filters.  How it is computed here (DESIGN.md has the derivation):
//
//   * every contig is packed to 2 bits per base (A0 T1 C2 G3, khmer's code) + 1 invalid bit, each
//     contig word-aligned in one arena (k_encode);
//   * genome 1 alone needs the full sketch: 2 bits per bin (seen once / seen again) updated with
//     L2-resident atomicOr (k_scan<COUNT1>), strand-specific polluters (windows touching a '~'
//     junction or a non-ACGT byte) go to two side bitmaps (k_junction, COUNT1);
//   * a genome-1 k-mer that survives the sketch owns its bin, so (k, bin) is a perfect hash of the
//     key set: a bitmap over bins + popcount ranks gives a dense slot id (k_scan<SELECT1>, k_rank,
//     k_scan<FILL1>) -- this replaces both the string sets and the Aho-Corasick automaton;
//   * every other genome is ONE pass (k_scan<SCAN>): window -> bin -> bitmap probe (L2); on a hit the
//     64-bit canonical word is compared with the slot's key: equal = an occurrence (position and
//     orientation OR-ed into the genome's row, a second one raises DUP), different = a sketch
//     collision in that genome (POLL).  A slot is shared iff every genome has exactly one
//     occurrence and no collision on the strand table the reference would have read;
//   * survivors are emitted in the reference's dict order by an ordered compaction over genome 1
//     (k_scan<EMIT_*>), filters are evaluated in fp64 with the reference's operation order, the
//     per-position "first k-mer that passes" pick is an atomicMin over a direct-mapped array.
//
// No tensor cores: nothing here is a contraction.  The hot kernel (SCAN) is bound by L2 sector
// reads (one 32 B sector per window and k), see DESIGN.md / profiles/.

#include "../../include/pfb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// row word layout (one uint32 per genome and slot)
constexpr uint32_t ROW_POS_MASK = 0x03FFFFFFu;  // leftmost start, genome-local padded coordinate
constexpr uint32_t ROW_ORI = 1u << 26;          // forward word is the canonical one (h <= r)
constexpr uint32_t ROW_HAS = 1u << 27;          // one occurrence seen
constexpr uint32_t ROW_DUP = 1u << 28;          // a second occurrence seen
constexpr uint32_t ROW_POLL = 1u << 29;         // a different k-mer of this genome shares the bin
constexpr uint32_t ROW_JF = 1u << 30;           // forward-string-only polluter ('~' or non-ACGT window)
constexpr uint32_t ROW_JR = 1u << 31;           // reverse-string-only polluter
constexpr uint32_t MAX_GENOME_WORDS = 1u << 21; // 2^26 padded bases

enum Mode { COUNT1 = 0, SELECT1 = 1, SCAN = 2, FILL1 = 3, EMIT_COUNT = 4, EMIT_WRITE = 5 };

struct K {  // kernel arguments (by value)
    int kmin, kmax, nk;
    uint32_t P, PW, SW;       // table prime, words per 1-bit bitmap, words per 2-bit sketch
    uint64_t barrett;         // floor(2^64 / P)
    int max_repeat, prefilter, dbg_counts;
    double min_gc, max_gc, min_tm, max_tm;
    double ln_salt, ln_dna4, ln_dna1, dmso_term, formamide;
    // arena
    const uint8_t *raw;
    uint64_t *seq2;
    uint32_t *nmask;
    uint32_t *wordContig;
    const uint32_t *cWordStart, *cLen, *cGenome, *cVirt;
    const uint64_t *cRawStart;
    const uint32_t *gWordStart, *gFirstContig, *gNumContigs, *gVirtLen;
    uint32_t nContigs, nGenomes, totalWords;
    // genome-1 sketch + key table
    uint32_t *sk, *jf, *jr, *bits, *rank, *slotBase;
    uint32_t *g1flags;        // [0] != 0: genome 1 has strand-specific polluters
    unsigned long long *allowedCnt;
    uint32_t *passmask1;
    uint64_t *canon1;
    uint8_t *slotK;
    uint32_t *rows;           // [nGenomes][nslots]
    uint64_t nslots;
    uint8_t *alive;
    // output
    uint32_t *emitCnt;        // per genome-1 word
    pfb_record *rec;
    uint32_t *recCount;
    pfb_pos *pos;             // [nGenomes][nrec]
    uint32_t *minRank;        // [totalWords*32]
    uint8_t *pick;
    uint64_t nrec;
};

// SantaLucia (1998) nearest-neighbour table in oligotm's integer units; index = 4*first + second
// with A0 T1 C2 G3; packed (S << 16) | H
__constant__ uint32_t c_nn[16] = {
    (222u << 16) | 79u,  (204u << 16) | 72u, (224u << 16) | 84u, (210u << 16) | 78u,
    (213u << 16) | 72u,  (222u << 16) | 79u, (222u << 16) | 82u, (227u << 16) | 85u,
    (227u << 16) | 85u,  (210u << 16) | 78u, (199u << 16) | 80u, (272u << 16) | 106u,
    (222u << 16) | 82u,  (224u << 16) | 84u, (244u << 16) | 98u, (199u << 16) | 80u};

__device__ __forceinline__ uint64_t kmask2(int k) { return ~0ull >> (64 - 2 * k); }
__device__ __forceinline__ uint64_t kmask1(int k) { return 0xFFFFFFFFull >> (32 - k); }

// reverse the 2-bit groups of a 64-bit word and complement (code ^ 1)
__device__ __forceinline__ uint64_t rc64(uint64_t x) {
    uint64_t y = __brevll(x);
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
    return y ^ 0x5555555555555555ull;
}

__device__ __forceinline__ uint32_t mod_p(uint64_t x, const K &p) {
    uint64_t q = __umul64hi(x, p.barrett);
    uint32_t r = (uint32_t)x - (uint32_t)q * p.P;
    return r >= p.P ? r - p.P : r;
}

// GC / repeat / Tm of one k-mer given as forward word (first base most significant).
// Operation order mirrors primer3's oligotm() as restated in oracle/pf_oracle.c so that the doubles
// agree bit for bit (no FMA contraction: explicit _rn intrinsics).
__device__ __forceinline__ int gc_count(uint64_t h) { return __popcll(h & 0xAAAAAAAAAAAAAAAAull); }

__device__ double oligo_tm(uint64_t h, int k, const K &p) {
    int dh = 0, ds = 0;
    uint64_t rc = rc64(h) >> (64 - 2 * k);
    bool sym = ((k & 1) == 0) && rc == h;
    if (sym) ds += 14;
    int first = (int)((h >> (2 * k - 2)) & 3), last = (int)(h & 3);
    if (first < 2) { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    if (last < 2)  { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    for (int i = 0; i < k - 1; i++) {
        uint32_t v = c_nn[(h >> (2 * i)) & 15];
        dh += (int)(v & 0xFFFF);
        ds += (int)(v >> 16);
    }
    int gc = gc_count(h);
    double delta_h = __dmul_rn((double)dh, -100.0);
    double delta_s = __dmul_rn((double)ds, -0.1);
    delta_s = __dadd_rn(delta_s, __dmul_rn(__dmul_rn(0.368, (double)(k - 1)), p.ln_salt));
    double ln_dna = sym ? p.ln_dna1 : p.ln_dna4;
    double tm = __dsub_rn(__ddiv_rn(delta_h, __dadd_rn(delta_s, __dmul_rn(1.987, ln_dna))), 273.15);
    tm = __dsub_rn(tm, p.dmso_term);
    tm = __dadd_rn(tm, __dmul_rn(__dsub_rn(__dmul_rn(0.453, __ddiv_rn((double)gc, (double)k)), 2.88), p.formamide));
    return tm;
}

__device__ __forceinline__ bool has_long_repeat(uint64_t h, int k, int R) {
    if (R <= 1) return true;
    if (k < R) return false;
    uint64_t e = h ^ (h >> 2);
    uint64_t z = ~(e | (e >> 1)) & 0x5555555555555555ull & kmask2(k - 1);  // pair (i, i+1) equal
    uint64_t y = z;
    for (int t = 1; t < R - 1; t++) y &= z >> (2 * t);
    return y != 0;
}

__device__ __forceinline__ uint32_t filter_flags(uint64_t h, int k, const K &p, double *tm_out) {
    uint32_t f = 0;
    double gcper = __dmul_rn(__ddiv_rn((double)gc_count(h), (double)k), 100.0);   // Primer.py:100
    if (gcper >= p.min_gc && gcper <= p.max_gc) f |= PFB_F_GC_OK;
    if (!has_long_repeat(h, k, p.max_repeat)) f |= PFB_F_REP_OK;
    double tm = oligo_tm(h, k, p);
    if (tm >= p.min_tm && tm <= p.max_tm) f |= PFB_F_TM_OK;
    if (tm_out) *tm_out = tm;
    return f;
}

// slot id of a set bit
__device__ __forceinline__ uint32_t slot_of(const K &p, int ki, uint32_t bin, uint32_t word) {
    return p.slotBase[ki] + p.rank[(size_t)ki * p.PW + (bin >> 5)] + __popc(word & ((1u << (bin & 31)) - 1u));
}

// ------------------------------------------------------------------------------------------
// k_encode: ASCII -> 2 bits / base (+ invalid mask); one thread per 32-base word.
// Replaces str(rec.seq) / reverse_complement / '~'.join (getCandidateKmers.py:93-100).
__global__ void __launch_bounds__(256) k_encode(K p, uint32_t w0, uint32_t w1) {
    uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= w1) return;
    // owning contig: last contig whose first word is <= w
    uint32_t lo = 0, hi = p.nContigs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.cWordStart[mid] <= w) lo = mid; else hi = mid;
    }
    uint32_t c = lo;
    uint32_t i = w - p.cWordStart[c];
    uint32_t len = p.cLen[c];
    uint32_t nvalid = min(32u, len - 32u * i);
    const uint8_t *src = p.raw + p.cRawStart[c] + 32ull * i;
    uintptr_t a = (uintptr_t)src;
    const uint32_t *src4 = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t x[9];
#pragma unroll
    for (int t = 0; t < 9; t++) x[t] = __ldg(src4 + t);   // arena has 64 B of slack
    uint64_t packed = 0;
    uint32_t inval = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        uint32_t v = sh ? __funnelshift_r(x[t], x[t + 1], sh) : x[t];
        uint32_t vm = __vcmpeq4(v, 0x41414141u) | __vcmpeq4(v, 0x43434343u) |
                      __vcmpeq4(v, 0x47474747u) | __vcmpeq4(v, 0x54545454u);   // 0xFF per valid byte
        uint32_t tt = (v >> 1) & 0x03030303u;                                  // A0 C1 T2 G3
        uint32_t code = ((tt & 0x01010101u) << 1) | ((tt >> 1) & 0x01010101u); // A0 T1 C2 G3
        code |= ~vm & 0x03030303u;                                            // everything else -> 3
        uint32_t pk = (code * 0x40100401u) >> 24;                              // byte 0 -> top 2 bits
        uint32_t nb = ((~vm & 0x01010101u) * 0x08040201u) >> 24;               // byte 0 -> bit 3
        packed |= (uint64_t)pk << (56 - 8 * t);
        inval |= (nb & 0xFu) << (28 - 4 * t);
    }
    if (nvalid < 32) {   // tail of the contig: pad = invalid
        uint32_t padmask = 0xFFFFFFFFu >> nvalid;
        inval |= padmask;
        packed |= ~0ull >> (2 * nvalid);
    }
    p.seq2[w] = packed;
    p.nmask[w] = inval;
    p.wordContig[w] = c;
    if (p.cGenome[c] == 0) {
        uint32_t real = nvalid < 32 ? inval & ~(0xFFFFFFFFu >> nvalid) : inval;
        if (real) atomicOr(&p.g1flags[0], 1u);
    }
}

// ------------------------------------------------------------------------------------------
// polluter bookkeeping shared by the scan kernels and k_junction
template <bool FIRST>
__device__ __forceinline__ void pollute(const K &p, int ki, uint32_t g, uint64_t canon, bool rev_table) {
    uint32_t bin = mod_p(canon, p);
    if (FIRST) {
        uint32_t *bm = rev_table ? p.jr : p.jf;
        atomicOr(&bm[(size_t)ki * p.PW + (bin >> 5)], 1u << (bin & 31));
    } else {
        uint32_t word = p.bits[(size_t)ki * p.PW + (bin >> 5)];
        if ((word >> (bin & 31)) & 1u) {
            uint32_t idx = slot_of(p, ki, bin, word);
            atomicOr(&p.rows[(size_t)g * p.nslots + idx], rev_table ? ROW_JR : ROW_JF);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_scan: one thread per 32-base word = 32 window end positions x nk lengths.
//   COUNT1   genome 1: sketch update (fkh.consume / rkh.consume, getCandidateKmers.py:46-51)
//   SELECT1  genome 1: the set comprehension + F - R (:54-59) (+ GC / repeat / Tm when prefilter)
//   FILL1    genome 1: write key / position of every slot
//   SCAN     other genomes: F ^ R membership (:54-55, 62-63), the per-k intersection (:127-142) and
//            the relocation (:160-187) in one probe
//   EMIT_*   ordered compaction of the surviving slots in the reference's dict order
template <int MODE>
__global__ void __launch_bounds__(256) k_scan(K p, uint32_t w0, uint32_t w1) {
    uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= w1) return;
    uint32_t c = p.wordContig[w];
    uint32_t cws = p.cWordStart[c], clen = p.cLen[c], g = p.cGenome[c];
    uint32_t i = w - cws;
    uint64_t cur = p.seq2[w], prev = i ? p.seq2[w - 1] : 0ull;
    uint64_t nm = ((uint64_t)(i ? p.nmask[w - 1] : 0u) << 32) | p.nmask[w];
    uint64_t hf = prev, hr = rc64(prev);
    uint32_t qbase = 32u * i;
    int nj = (int)min(32u, clen - qbase);
    uint32_t gpos0 = 32u * (w - p.gWordStart[g]);   // genome-local padded coordinate of base j = 0
    const bool g1poll = (MODE == SELECT1) ? (p.g1flags[0] != 0) : false;
    uint32_t emitted = 0;
    uint32_t emit_at = 0;
    if (MODE == EMIT_WRITE) emit_at = p.emitCnt[w - w0];

    for (int j = 0; j < nj; j++) {
        uint32_t code = (uint32_t)(cur >> (62 - 2 * j)) & 3u;
        hf = (hf << 2) | code;
        hr = (hr >> 2) | ((uint64_t)(code ^ 1u) << 62);
        int n = (int)qbase + j + 1;                 // bases of this contig up to and including j
        int khi = min(p.kmax, n);
        uint32_t pm = 0;
        if (MODE == FILL1 || MODE == EMIT_COUNT || MODE == EMIT_WRITE) {
            pm = p.passmask1[gpos0 + j];
            if (!pm) continue;
        }
        const uint64_t nmw = nm >> (31 - j);

        if (MODE == EMIT_COUNT || MODE == EMIT_WRITE || MODE == FILL1) {
            // longest first at one end position (pyahocorasick emission order)
            for (int k = khi; k >= p.kmin; k--) {
                int ki = k - p.kmin;
                if (!((pm >> ki) & 1u)) continue;
                uint64_t h = hf & kmask2(k), r = hr >> (64 - 2 * k);
                uint64_t canon = h < r ? h : r;
                uint32_t bin = mod_p(canon, p);
                uint32_t word = p.bits[(size_t)ki * p.PW + (bin >> 5)];
                uint32_t idx = slot_of(p, ki, bin, word);
                if (MODE == FILL1) {
                    p.canon1[idx] = canon;
                    p.slotK[idx] = (uint8_t)k;
                    p.rows[idx] = ROW_HAS | (h <= r ? ROW_ORI : 0u) | (gpos0 + j - k + 1);
                } else {
                    if (!p.alive[idx]) continue;
                    if (MODE == EMIT_WRITE) {
                        pfb_record rec;
                        double tm;
                        uint32_t f = filter_flags(h, k, p, &tm);
                        if (h == r) f |= PFB_F_PALIN;
                        rec.word = h; rec.tm = tm; rec.slot = idx; rec.gc_count = (uint16_t)gc_count(h);
                        rec.k = (uint8_t)k; rec.flags = (uint8_t)f;
                        p.rec[emit_at + emitted] = rec;
                    }
                    emitted++;
                }
            }
            continue;
        }

        for (int k = p.kmin; k <= khi; k++) {
            int ki = k - p.kmin;
            uint64_t h = hf & kmask2(k), r = hr >> (64 - 2 * k);
            uint32_t wn = (uint32_t)(nmw & kmask1(k));      // invalid bases inside the window (bit t = age t)
            if (wn) {
                if (MODE == SELECT1) continue;
                // strand-specific polluter: the forward string hashes the byte as (3, comp 2); the
                // reverse-complement string keeps the byte, so there it is again (3, comp 2)
                uint64_t canon_f = h < r ? h : r;
                uint64_t h2 = r, r2 = h;
                for (int t = 0; t < k; t++)
                    if ((wn >> t) & 1u) { h2 |= 1ull << (2 * (k - 1 - t)); r2 &= ~(1ull << (2 * t)); }
                uint64_t canon_r = h2 < r2 ? h2 : r2;
                pollute<MODE == COUNT1>(p, ki, g, canon_f, false);
                pollute<MODE == COUNT1>(p, ki, g, canon_r, true);
                continue;
            }
            uint64_t canon = h < r ? h : r;
            if (MODE == SELECT1) {
                // isOneEndGc (:29-33) before touching memory
                if (((h >> (2 * k - 2)) & 3) < 2 && (h & 3) < 2) continue;
            }
            uint32_t bin = mod_p(canon, p);
            if (MODE == COUNT1) {
                uint32_t *cell = &p.sk[(size_t)ki * p.SW + (bin >> 4)];
                uint32_t b = bin & 15u;
                uint32_t old = atomicOr(cell, 1u << b);
                if ((old >> b) & 1u) atomicOr(cell, 1u << (16 + b));
            } else if (MODE == SELECT1) {
                uint32_t st = p.sk[(size_t)ki * p.SW + (bin >> 4)];
                uint32_t b = bin & 15u;
                if (!((st >> b) & 1u) || ((st >> (16 + b)) & 1u)) continue;   // get(x) == 1 (:35-38)
                bool inF = true, inR = true;
                if (g1poll) {
                    inF = !((p.jf[(size_t)ki * p.PW + (bin >> 5)] >> (bin & 31)) & 1u);
                    inR = !((p.jr[(size_t)ki * p.PW + (bin >> 5)] >> (bin & 31)) & 1u);
                }
                bool keep = (h == r) ? (inF && !inR) : inF;   // first genome: F - R (:58-59)
                if (!keep) continue;
                if (p.dbg_counts) atomicAdd(&p.allowedCnt[ki], 1ull);
                if (p.prefilter && (filter_flags(h, k, p, nullptr) & 7u) != 7u) continue;
                atomicOr(&p.bits[(size_t)ki * p.PW + (bin >> 5)], 1u << (bin & 31));
                pm |= 1u << ki;
            } else if (MODE == SCAN) {
                uint32_t word = __ldg(&p.bits[(size_t)ki * p.PW + (bin >> 5)]);
                if ((word >> (bin & 31)) & 1u) {
                    uint32_t idx = slot_of(p, ki, bin, word);
                    uint32_t *row = &p.rows[(size_t)g * p.nslots + idx];
                    if (p.canon1[idx] == canon) {
                        uint32_t val = ROW_HAS | (h <= r ? ROW_ORI : 0u) | (gpos0 + j - k + 1);
                        uint32_t old = atomicOr(row, val);
                        if (old & ROW_HAS) atomicOr(row, ROW_DUP);
                    } else {
                        atomicOr(row, ROW_POLL);
                    }
                }
            }
        }
        if (MODE == SELECT1) p.passmask1[gpos0 + j] = pm;
    }
    if (MODE == EMIT_COUNT) p.emitCnt[w - w0] = emitted;
}

// ------------------------------------------------------------------------------------------
// k_junction: windows of the '~'-joined strings that touch a junction (getCandidateKmers.py:99-100):
// they never become k-mers (:40-43) but khmer counts them ('~' hashes like G).  One thread per
// (contig boundary, k, strand string, window offset) in the virtual joined coordinates.
template <bool FIRST>
__global__ void __launch_bounds__(128) k_junction(K p, uint32_t c0, uint32_t c1) {
    const int T = 2 * p.kmax - 1;
    unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t = (int)(tid % T); tid /= T;
    int rev = (int)(tid & 1); tid >>= 1;
    int ki = (int)(tid % p.nk); tid /= p.nk;
    if (tid >= (unsigned long long)(c1 - c0)) return;
    uint32_t c = c0 + (uint32_t)tid;
    if (c + 1 >= p.nContigs || p.cGenome[c + 1] != p.cGenome[c]) return;   // not a junction
    uint32_t g = p.cGenome[c];
    int k = p.kmin + ki;
    if (t >= p.kmax + k - 1) return;
    long long jstart = (long long)p.cVirt[c] + p.cLen[c];
    long long jend = jstart + p.kmax - 1;
    long long s = jstart - k + 1 + t;
    if (s < (long long)p.cVirt[c] || s > jend) return;          // belongs to an earlier junction / past this one
    if (s + k > (long long)p.gVirtLen[g]) return;
    uint32_t gc_last = p.gFirstContig[g] + p.gNumContigs[g] - 1;
    uint32_t ci = c;
    uint64_t h = 0, r = 0;
    for (int q = 0; q < k; q++) {
        long long v = s + q;
        while (ci < gc_last && v >= (long long)p.cVirt[ci + 1]) ci++;
        long long off = v - (long long)p.cVirt[ci];
        uint32_t cf = 3, cc = 2;
        if (off < (long long)p.cLen[ci]) {
            uint32_t o = rev ? (p.cLen[ci] - 1 - (uint32_t)off) : (uint32_t)off;
            uint32_t wd = p.cWordStart[ci] + (o >> 5);
            uint32_t b = (uint32_t)(p.seq2[wd] >> (62 - 2 * (o & 31))) & 3u;
            bool bad = (p.nmask[wd] >> (31 - (o & 31))) & 1u;
            if (!bad) { cf = rev ? (b ^ 1u) : b; cc = cf ^ 1u; }
        }
        h = (h << 2) | cf;
        r = (r >> 2) | ((uint64_t)cc << (2 * (k - 1)));
    }
    uint64_t canon = h < r ? h : r;
    if (FIRST) atomicOr(&p.g1flags[0], 1u);
    pollute<FIRST>(p, ki, g, canon, rev != 0);
}

// ------------------------------------------------------------------------------------------
// k_rank: exclusive popcount prefix of each k's bitmap (one block per k) -> dense slot ids
__global__ void __launch_bounds__(1024) k_rank(K p, uint32_t *ktotal) {
    __shared__ uint32_t sm[1024];
    int ki = blockIdx.x;
    const uint32_t *bm = p.bits + (size_t)ki * p.PW;
    uint32_t *rk = p.rank + (size_t)ki * p.PW;
    uint32_t per = (p.PW + blockDim.x - 1) / blockDim.x;
    uint32_t a = min(p.PW, threadIdx.x * per), b = min(p.PW, a + per);
    uint32_t s = 0;
    for (uint32_t x = a; x < b; x++) s += __popc(bm[x]);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= off ? sm[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sm[threadIdx.x] - s;
    for (uint32_t x = a; x < b; x++) { rk[x] = run; run += __popc(bm[x]); }
    if (threadIdx.x == 1023) ktotal[ki] = sm[1023];
}

__global__ void k_slotbase(K p, const uint32_t *ktotal) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t s = 0;
        for (int ki = 0; ki < p.nk; ki++) { p.slotBase[ki] = s; s += ktotal[ki]; }
        p.slotBase[p.nk] = s;
    }
}

// exclusive scan of a uint32 array, single block (n up to a few million; used on per-word counts)
__global__ void __launch_bounds__(1024) k_exscan(uint32_t *a, uint32_t n, uint32_t *total) {
    __shared__ uint32_t sm[1024];
    uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    uint32_t s = 0;
    for (uint32_t x = lo; x < hi; x++) s += a[x];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= off ? sm[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sm[threadIdx.x] - s;
    for (uint32_t x = lo; x < hi; x++) { uint32_t v = a[x]; a[x] = run; run += v; }
    if (threadIdx.x == 1023) *total = sm[1023];
}

// ------------------------------------------------------------------------------------------
// k_alive: per slot, the reference's rules for every other genome on this context.
//   occurrence on the forward string ('+'): needs count == 1 in fkh  -> no JF polluter
//   occurrence on the rc string ('-')     : needs count == 1 in rkh  -> no JR polluter
//   palindrome: in F ^ R  <=>  exactly one of the two tables reads 1 (:62-63)
__device__ __forceinline__ bool row_ok(uint32_t v, bool ori1, bool pal) {
    if (!(v & ROW_HAS) || (v & (ROW_DUP | ROW_POLL))) return false;
    if (pal) return ((v & ROW_JF) != 0) != ((v & ROW_JR) != 0);
    bool minus = ((v & ROW_ORI) != 0) != ori1;
    return minus ? !(v & ROW_JR) : !(v & ROW_JF);
}

__global__ void __launch_bounds__(256) k_alive(K p) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.nslots) return;
    uint64_t canon = p.canon1[idx];
    int k = p.slotK[idx];
    bool pal = (rc64(canon) >> (64 - 2 * k)) == canon;
    bool ori1 = (p.rows[idx] & ROW_ORI) != 0;
    bool ok = true;
    for (uint32_t g = 1; g < p.nGenomes; g++) ok = ok && row_ok(p.rows[(size_t)g * p.nslots + idx], ori1, pal);
    p.alive[idx] = ok ? 1 : 0;
}

// k_positions: (contig, leftmost start, strand) of every record in every genome
// (getCandidateKmers.py:173-187); also seeds the grouping array with the smallest passing rank
__global__ void __launch_bounds__(256) k_positions(K p) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t g = blockIdx.y;
    if (r >= p.nrec) return;
    pfb_record rec = p.rec[r];
    uint32_t v = p.rows[(size_t)g * p.nslots + rec.slot];
    bool ori1 = (p.rows[rec.slot] & ROW_ORI) != 0;
    bool pal = (rec.flags & PFB_F_PALIN) != 0;
    uint32_t lp = v & ROW_POS_MASK;
    uint32_t gw = p.gWordStart[g] + (lp >> 5);
    uint32_t c = p.wordContig[gw];
    pfb_pos o;
    o.contig = c - p.gFirstContig[g];
    o.start = lp - 32u * (p.cWordStart[c] - p.gWordStart[g]);
    o.strand = pal ? 1u : ((((v & ROW_ORI) != 0) != ori1) ? 1u : 0u);   // later '-' match overwrites (:181-187)
    p.pos[(size_t)g * p.nrec + r] = o;
    if ((rec.flags & 7u) == 7u) atomicMin(&p.minRank[(size_t)p.gWordStart[g] * 32u + lp], (uint32_t)r);
}

// k_pick: a record is picked when it is the first passing k-mer of its (contig, start) group in
// some genome (getCandidateKmers.py:222-247, 376-388)
__global__ void __launch_bounds__(256) k_pick(K p) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t g = blockIdx.y;
    if (r >= p.nrec) return;
    pfb_record rec = p.rec[r];
    if ((rec.flags & 7u) != 7u) return;
    uint32_t v = p.rows[(size_t)g * p.nslots + rec.slot];
    uint32_t lp = v & ROW_POS_MASK;
    if (p.minRank[(size_t)p.gWordStart[g] * 32u + lp] == (uint32_t)r) p.pick[r] = 1;
}

__global__ void k_apply_pick(K p, unsigned long long *npick) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.nrec) return;
    if (p.pick[r]) { p.rec[r].flags |= PFB_F_PICK; atomicAdd(npick, 1ull); }
}

__global__ void k_oligo(K p, uint64_t h, int k, double *out) {
    out[0] = oligo_tm(h, k, p);
    out[1] = __dmul_rn(__ddiv_rn((double)gc_count(h), (double)k), 100.0);
}

// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct HostGenome {
    const uint8_t *bases;
    uint64_t n_bases;
    std::vector<uint64_t> off;   // n_contigs + 1
};

}  // namespace

struct pfb_ctx {
    int device = 0;
    int dbg_counts = 0;
    pfb_params prm{};
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8]{};
    std::vector<HostGenome> genomes;
    // layout (host)
    std::vector<uint32_t> cWordStart, cLen, cGenome, cVirt, gWordStart, gFirstContig, gNumContigs, gVirtLen;
    std::vector<uint64_t> cRawStart, gRawStart;
    uint64_t totalRaw = 0;
    uint32_t totalWords = 0;
    bool uploaded = false, computed = false;
    int stage_done = -1;
    // device
    DevBuf raw, seq2, nmask, wordContig, tabs32, tabs64, sk, jf, jr, bits, rank, misc, passmask1, canon1, slotK,
        rows, alive, emitCnt, rec, pos, minRank, pick;
    K k{};
    uint64_t nslots = 0, nrec = 0, npick = 0;
    uint64_t launches = 0;
    pfb_timing tm{};
    std::vector<uint64_t> allowed;
};

static std::string g_create_err;

namespace {

int fail(pfb_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, PFB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

int ensure(pfb_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return PFB_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return PFB_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

inline unsigned nblk(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int validate(const pfb_params *p, std::string &why) {
    if (!p) { why = "params is NULL"; return PFB_EINVAL; }
    if (p->k_min < 1 || p->k_max > 32 || p->k_min > p->k_max) { why = "need 1 <= k_min <= k_max <= 32"; return PFB_EINVAL; }
    if (p->table_size < 64) { why = "table_size too small"; return PFB_EINVAL; }
    if (p->max_repeat < 1) { why = "max_repeat must be >= 1"; return PFB_EINVAL; }
    return PFB_OK;
}

// layout of the arena from the host-side genome list
int build_layout(pfb_ctx *ctx) {
    auto &G = ctx->genomes;
    ctx->cWordStart.clear(); ctx->cLen.clear(); ctx->cGenome.clear(); ctx->cVirt.clear(); ctx->cRawStart.clear();
    ctx->gWordStart.clear(); ctx->gFirstContig.clear(); ctx->gNumContigs.clear(); ctx->gVirtLen.clear(); ctx->gRawStart.clear();
    uint64_t raw = 0;
    uint64_t words = 0;
    for (size_t g = 0; g < G.size(); g++) {
        raw = (raw + 255) & ~255ull;
        ctx->gRawStart.push_back(raw);
        ctx->gWordStart.push_back((uint32_t)words);
        ctx->gFirstContig.push_back((uint32_t)ctx->cLen.size());
        uint32_t nc = (uint32_t)G[g].off.size() - 1;
        ctx->gNumContigs.push_back(nc);
        uint64_t virt = 0, gwords = 0;
        for (uint32_t c = 0; c < nc; c++) {
            uint64_t len = G[g].off[c + 1] - G[g].off[c];
            ctx->cWordStart.push_back((uint32_t)(words + gwords));
            ctx->cLen.push_back((uint32_t)len);
            ctx->cGenome.push_back((uint32_t)g);
            ctx->cVirt.push_back((uint32_t)virt);
            ctx->cRawStart.push_back(raw + G[g].off[c]);
            virt += len + (c + 1 < nc ? (uint64_t)ctx->prm.k_max : 0);
            gwords += (len + 31) / 32;
        }
        if (gwords > MAX_GENOME_WORDS || virt >= (1ull << 31))
            return fail(ctx, PFB_ETOOLARGE, "genome " + std::to_string(g) + " exceeds 2^26 padded bases");
        ctx->gVirtLen.push_back((uint32_t)virt);
        words += gwords;
        raw += G[g].n_bases;
        if (words >= (1ull << 31)) return fail(ctx, PFB_ETOOLARGE, "more than 2^36 bases on one context");
    }
    ctx->totalRaw = raw;
    ctx->totalWords = (uint32_t)words;
    return PFB_OK;
}

}  // namespace

// ==========================================================================================
extern "C" {

int pfb_default_params(pfb_params *p) {
    if (!p) return PFB_EINVAL;
    memset(p, 0, sizeof(*p));
    p->k_min = 16; p->k_max = 20; p->table_size = 9999991u; p->max_repeat = 4;
    p->min_gc = 40.0; p->max_gc = 60.0; p->min_tm = 55.0; p->max_tm = 68.0;
    p->mv_conc = 50.0; p->dv_conc = 1.5; p->dntp_conc = 0.6; p->dna_conc = 50.0;
    p->dmso_conc = 0.0; p->dmso_fact = 0.6; p->formamide_conc = 0.8;
    p->prefilter = 1;
    return PFB_OK;
}

const char *pfb_last_error(const pfb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pfb_create(pfb_ctx **out, int device, const pfb_params *params) {
    pfb_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, PFB_EINVAL, "out is NULL");
    *out = nullptr;
    std::string why;
    if (validate(params, why) != PFB_OK) return fail(nullptr, PFB_EINVAL, why);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PFB_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, PFB_EINVAL, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, PFB_ECUDA, cudaGetErrorString(e));
    ctx = new pfb_ctx();
    ctx->device = device;
    ctx->prm = *params;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete ctx; return fail(nullptr, PFB_ECUDA, m); }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    *out = ctx;
    return PFB_OK;
}

void pfb_destroy(pfb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *all[] = {&ctx->raw, &ctx->seq2, &ctx->nmask, &ctx->wordContig, &ctx->tabs32, &ctx->tabs64, &ctx->sk, &ctx->jf,
                     &ctx->jr, &ctx->bits, &ctx->rank, &ctx->misc, &ctx->passmask1, &ctx->canon1, &ctx->slotK, &ctx->rows,
                     &ctx->alive, &ctx->emitCnt, &ctx->rec, &ctx->pos, &ctx->minRank, &ctx->pick};
    for (auto *b : all) release(*b);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pfb_add_genome(pfb_ctx *ctx, const uint8_t *bases, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs) {
    if (!ctx) return PFB_EINVAL;
    if ((!bases && n_bases) || !contig_off || n_contigs == 0) return fail(ctx, PFB_EINVAL, "bad genome arguments");
    if (contig_off[0] != 0 || contig_off[n_contigs] != n_bases) return fail(ctx, PFB_EINVAL, "contig offsets must span [0, n_bases]");
    for (uint32_t c = 0; c < n_contigs; c++)
        if (contig_off[c + 1] < contig_off[c]) return fail(ctx, PFB_EINVAL, "contig offsets must be non-decreasing");
    HostGenome g;
    g.bases = bases; g.n_bases = n_bases;
    g.off.assign(contig_off, contig_off + n_contigs + 1);
    ctx->genomes.push_back(std::move(g));
    ctx->uploaded = false; ctx->computed = false; ctx->stage_done = -1;
    return PFB_OK;
}

int pfb_clear_genomes(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    ctx->genomes.clear();
    ctx->uploaded = false; ctx->computed = false; ctx->stage_done = -1;
    return PFB_OK;
}

int pfb_sync(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return PFB_OK;
}

int pfb_upload(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    if (ctx->genomes.empty()) return fail(ctx, PFB_ESTATE, "no genome added");
    CK(cudaSetDevice(ctx->device));
    int rc = build_layout(ctx);
    if (rc != PFB_OK) return rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ensure(ctx, ctx->raw, ctx->totalRaw + 64)) != PFB_OK) return rc;
    CK(cudaEventRecord(ctx->ev[0], st));
    for (size_t g = 0; g < ctx->genomes.size(); g++) {
        auto &hg = ctx->genomes[g];
        if (hg.n_bases)
            CK(cudaMemcpyAsync((uint8_t *)ctx->raw.p + ctx->gRawStart[g], hg.bases, hg.n_bases, cudaMemcpyHostToDevice, st));
    }
    // contig / genome tables
    size_t nC = ctx->cLen.size(), nG = ctx->genomes.size();
    std::vector<uint32_t> t32;
    auto put32 = [&](const std::vector<uint32_t> &v) { size_t o = t32.size(); t32.insert(t32.end(), v.begin(), v.end()); return o; };
    size_t o_cws = put32(ctx->cWordStart), o_cl = put32(ctx->cLen), o_cg = put32(ctx->cGenome), o_cv = put32(ctx->cVirt);
    size_t o_gws = put32(ctx->gWordStart), o_gfc = put32(ctx->gFirstContig), o_gnc = put32(ctx->gNumContigs), o_gvl = put32(ctx->gVirtLen);
    if ((rc = ensure(ctx, ctx->tabs32, t32.size() * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->tabs64, nC * 8)) != PFB_OK) return rc;
    CK(cudaMemcpyAsync(ctx->tabs32.p, t32.data(), t32.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->tabs64.p, ctx->cRawStart.data(), nC * 8, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaStreamSynchronize(st));   // t32 / host genome buffers are borrowed only until here
    const uint32_t *b32 = (const uint32_t *)ctx->tabs32.p;
    K &k = ctx->k;
    k.cWordStart = b32 + o_cws; k.cLen = b32 + o_cl; k.cGenome = b32 + o_cg; k.cVirt = b32 + o_cv;
    k.gWordStart = b32 + o_gws; k.gFirstContig = b32 + o_gfc; k.gNumContigs = b32 + o_gnc; k.gVirtLen = b32 + o_gvl;
    k.cRawStart = (const uint64_t *)ctx->tabs64.p;
    k.nContigs = (uint32_t)nC; k.nGenomes = (uint32_t)nG; k.totalWords = ctx->totalWords;
    k.raw = (const uint8_t *)ctx->raw.p;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tm.ms_upload = ms;
    ctx->uploaded = true; ctx->computed = false; ctx->stage_done = -1;
    return PFB_OK;
}

static int stage0(pfb_ctx *ctx) {
    const pfb_params &P = ctx->prm;
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    int rc;
    const int nk = P.k_max - P.k_min + 1;
    k.kmin = P.k_min; k.kmax = P.k_max; k.nk = nk;
    k.P = P.table_size; k.PW = (P.table_size + 31) / 32; k.SW = (P.table_size + 15) / 16;
    k.barrett = (uint64_t)((((unsigned __int128)1) << 64) / P.table_size);
    k.max_repeat = P.max_repeat; k.prefilter = P.prefilter; k.dbg_counts = ctx->dbg_counts;
    k.min_gc = P.min_gc; k.max_gc = P.max_gc; k.min_tm = P.min_tm; k.max_tm = P.max_tm;
    {   // constants of oligotm(), computed with the host libm exactly like the reference's C code does
        double dv = P.dv_conc, dntp = P.dntp_conc;
        if (dv == 0) dntp = 0;
        if (dv < dntp) dv = dntp;
        double k_mm = P.mv_conc + 120.0 * std::sqrt(dv - dntp);
        k.ln_salt = std::log(k_mm / 1000.0);
        k.ln_dna4 = std::log(P.dna_conc / 4000000000.0);
        k.ln_dna1 = std::log(P.dna_conc / 1000000000.0);
        k.dmso_term = P.dmso_conc * P.dmso_fact;
        k.formamide = P.formamide_conc;
    }
    const uint32_t W = ctx->totalWords;
    const uint32_t W1 = ctx->genomes.size() > 1 ? ctx->gWordStart[1] : W;   // genome 1 = words [0, W1)
    ctx->launches = 0;
    // buffers
    if ((rc = ensure(ctx, ctx->seq2, (size_t)(W + 1) * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->nmask, (size_t)(W + 1) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->wordContig, (size_t)(W + 1) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->sk, (size_t)nk * k.SW * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->jf, (size_t)nk * k.PW * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->jr, (size_t)nk * k.PW * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->bits, (size_t)nk * k.PW * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->rank, (size_t)nk * k.PW * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->misc, 4096)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->passmask1, (size_t)W1 * 32 * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->emitCnt, (size_t)(W1 + 1) * 4)) != PFB_OK) return rc;
    k.seq2 = (uint64_t *)ctx->seq2.p; k.nmask = (uint32_t *)ctx->nmask.p; k.wordContig = (uint32_t *)ctx->wordContig.p;
    k.sk = (uint32_t *)ctx->sk.p; k.jf = (uint32_t *)ctx->jf.p; k.jr = (uint32_t *)ctx->jr.p;
    k.bits = (uint32_t *)ctx->bits.p; k.rank = (uint32_t *)ctx->rank.p;
    // misc: [0..63] slotBase, [64] g1flags, [65] recCount, [66..67] npick, [128..191] ktotal, [256..] allowedCnt (u64)
    uint32_t *misc = (uint32_t *)ctx->misc.p;
    k.slotBase = misc; k.g1flags = misc + 64; k.recCount = misc + 65;
    uint32_t *ktotal = misc + 128;
    k.allowedCnt = (unsigned long long *)(misc + 256);
    k.passmask1 = (uint32_t *)ctx->passmask1.p; k.emitCnt = (uint32_t *)ctx->emitCnt.p;

    CK(cudaEventRecord(ctx->ev[2], st));
    CK(cudaMemsetAsync(ctx->misc.p, 0, 4096, st));
    CK(cudaMemsetAsync(ctx->sk.p, 0, (size_t)nk * k.SW * 4, st));
    CK(cudaMemsetAsync(ctx->jf.p, 0, (size_t)nk * k.PW * 4, st));
    CK(cudaMemsetAsync(ctx->jr.p, 0, (size_t)nk * k.PW * 4, st));
    CK(cudaMemsetAsync(ctx->bits.p, 0, (size_t)nk * k.PW * 4, st));
    if (ctx->gNumContigs[0] > 1) {
        uint32_t one = 1;
        CK(cudaMemcpyAsync(k.g1flags, &one, 4, cudaMemcpyHostToDevice, st));
    }
    // encode everything
    if (W) { k_encode<<<nblk(W, 256), 256, 0, st>>>(k, 0, W); ctx->launches++; }
    CK(cudaEventRecord(ctx->ev[3], st));
    // genome 1
    if (W1) { k_scan<COUNT1><<<nblk(W1, 256), 256, 0, st>>>(k, 0, W1); ctx->launches++; }
    if (ctx->gNumContigs[0] > 1) {
        uint64_t nthr = (uint64_t)ctx->gNumContigs[0] * nk * 2 * (2 * P.k_max - 1);
        k_junction<true><<<nblk(nthr, 128), 128, 0, st>>>(k, 0, ctx->gNumContigs[0]);
        ctx->launches++;
    }
    if (W1) { k_scan<SELECT1><<<nblk(W1, 256), 256, 0, st>>>(k, 0, W1); ctx->launches++; }
    k_rank<<<nk, 1024, 0, st>>>(k, ktotal); ctx->launches++;
    k_slotbase<<<1, 32, 0, st>>>(k, ktotal); ctx->launches++;
    uint32_t hmisc[64 + 1];
    CK(cudaMemcpyAsync(hmisc, misc, sizeof(hmisc), cudaMemcpyDeviceToHost, st));
    ctx->allowed.assign(nk, 0);
    CK(cudaMemcpyAsync(ctx->allowed.data(), k.allowedCnt, (size_t)nk * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->nslots = hmisc[nk];
    k.nslots = ctx->nslots;
    const size_t NS = ctx->nslots ? ctx->nslots : 1;
    const size_t nG = ctx->genomes.size();
    if ((rc = ensure(ctx, ctx->canon1, NS * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->slotK, NS)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->rows, NS * nG * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->alive, NS)) != PFB_OK) return rc;
    k.canon1 = (uint64_t *)ctx->canon1.p; k.slotK = (uint8_t *)ctx->slotK.p;
    k.rows = (uint32_t *)ctx->rows.p; k.alive = (uint8_t *)ctx->alive.p;
    CK(cudaMemsetAsync(ctx->rows.p, 0, NS * nG * 4, st));
    if (W1 && ctx->nslots) { k_scan<FILL1><<<nblk(W1, 256), 256, 0, st>>>(k, 0, W1); ctx->launches++; }
    CK(cudaEventRecord(ctx->ev[4], st));
    // every other genome: one pass against the key table
    if (W > W1 && ctx->nslots) {
        k_scan<SCAN><<<nblk(W - W1, 256), 256, 0, st>>>(k, W1, W); ctx->launches++;
        uint32_t c0 = ctx->gFirstContig[1], c1 = (uint32_t)ctx->cLen.size();
        if (c1 - c0 > nG - 1) {   // some genome has more than one contig
            uint64_t nthr = (uint64_t)(c1 - c0) * nk * 2 * (2 * P.k_max - 1);
            k_junction<false><<<nblk(nthr, 128), 128, 0, st>>>(k, c0, c1);
            ctx->launches++;
        }
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    if (ctx->nslots) { k_alive<<<nblk(ctx->nslots, 256), 256, 0, st>>>(k); ctx->launches++; }
    CK(cudaGetLastError());
    return PFB_OK;
}

static int stage1(pfb_ctx *ctx) {
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    int rc;
    const uint32_t W = ctx->totalWords;
    const uint32_t W1 = ctx->genomes.size() > 1 ? ctx->gWordStart[1] : W;
    const size_t nG = ctx->genomes.size();
    ctx->nrec = 0; ctx->npick = 0;
    if (ctx->nslots && W1) {
        k_scan<EMIT_COUNT><<<nblk(W1, 256), 256, 0, st>>>(k, 0, W1); ctx->launches++;
        k_exscan<<<1, 1024, 0, st>>>(k.emitCnt, W1, k.recCount); ctx->launches++;
        uint32_t n = 0;
        CK(cudaMemcpyAsync(&n, k.recCount, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->nrec = n;
    }
    k.nrec = ctx->nrec;
    const size_t NR = ctx->nrec ? ctx->nrec : 1;
    if ((rc = ensure(ctx, ctx->rec, NR * sizeof(pfb_record))) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->pos, NR * nG * sizeof(pfb_pos))) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->pick, NR)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->minRank, (size_t)W * 32 * 4 + 4)) != PFB_OK) return rc;
    k.rec = (pfb_record *)ctx->rec.p; k.pos = (pfb_pos *)ctx->pos.p; k.pick = (uint8_t *)ctx->pick.p;
    k.minRank = (uint32_t *)ctx->minRank.p;
    CK(cudaMemsetAsync(ctx->pick.p, 0, NR, st));
    if (ctx->nrec) {
        CK(cudaMemsetAsync(ctx->minRank.p, 0xFF, (size_t)W * 32 * 4, st));
        k_scan<EMIT_WRITE><<<nblk(W1, 256), 256, 0, st>>>(k, 0, W1); ctx->launches++;
        dim3 grid(nblk(ctx->nrec, 256), (unsigned)nG);
        k_positions<<<grid, 256, 0, st>>>(k); ctx->launches++;
        k_pick<<<grid, 256, 0, st>>>(k); ctx->launches++;
    }
    CK(cudaGetLastError());
    return PFB_OK;
}

static int stage2(pfb_ctx *ctx) {
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    unsigned long long *npick = (unsigned long long *)(k.recCount + 1);
    if (ctx->nrec) {
        k_apply_pick<<<nblk(ctx->nrec, 256), 256, 0, st>>>(k, npick); ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[6], st));
    unsigned long long np = 0;
    CK(cudaMemcpyAsync(&np, npick, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    ctx->npick = np;
    pfb_timing &t = ctx->tm;
    cudaEventElapsedTime(&t.ms_encode, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&t.ms_first, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&t.ms_scan, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&t.ms_finalize, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&t.ms_total, ctx->ev[2], ctx->ev[6]);
    uint64_t nb = 0, nw = 0;
    const int nk = ctx->prm.k_max - ctx->prm.k_min + 1;
    for (auto &g : ctx->genomes) {
        nb += g.n_bases;
        for (size_t c = 0; c + 1 < g.off.size(); c++) {
            uint64_t len = g.off[c + 1] - g.off[c];
            for (int kk = ctx->prm.k_min; kk <= ctx->prm.k_max; kk++) if (len >= (uint64_t)kk) nw += len - kk + 1;
        }
    }
    (void)nk;
    t.n_bases = nb; t.n_windows = nw; t.n_slots = ctx->nslots; t.n_records = ctx->nrec; t.n_picked = ctx->npick;
    t.n_launches = ctx->launches;
    ctx->computed = true;
    return PFB_OK;
}

int pfb_compute_stage(pfb_ctx *ctx, int stage) {
    if (!ctx) return PFB_EINVAL;
    if (!ctx->uploaded) return fail(ctx, PFB_ESTATE, "pfb_upload has not run");
    if (stage < 0 || stage > 2) return fail(ctx, PFB_EINVAL, "stage must be 0, 1 or 2");
    if (stage > 0 && ctx->stage_done < stage - 1) return fail(ctx, PFB_ESTATE, "stages must run in order");
    CK(cudaSetDevice(ctx->device));
    int rc = stage == 0 ? stage0(ctx) : stage == 1 ? stage1(ctx) : stage2(ctx);
    if (rc == PFB_OK) ctx->stage_done = stage;
    return rc;
}

int pfb_compute(pfb_ctx *ctx) {
    int rc;
    for (int s = 0; s < 3; s++)
        if ((rc = pfb_compute_stage(ctx, s)) != PFB_OK) return rc;
    return PFB_OK;
}

int pfb_run(pfb_ctx *ctx) {
    int rc = pfb_upload(ctx);
    if (rc != PFB_OK) return rc;
    return pfb_compute(ctx);
}

int pfb_device_buffer(pfb_ctx *ctx, int which, void **dev_ptr, uint64_t *n_bytes) {
    if (!ctx || !dev_ptr || !n_bytes) return PFB_EINVAL;
    if (which == PFB_BUF_ALIVE) {
        if (ctx->stage_done < 0) return fail(ctx, PFB_ESTATE, "stage 0 has not run");
        *dev_ptr = ctx->alive.p; *n_bytes = ctx->nslots;
    } else if (which == PFB_BUF_PICK) {
        if (ctx->stage_done < 1) return fail(ctx, PFB_ESTATE, "stage 1 has not run");
        *dev_ptr = ctx->pick.p; *n_bytes = ctx->nrec;
    } else {
        return fail(ctx, PFB_EINVAL, "unknown buffer");
    }
    return PFB_OK;
}

int pfb_result_counts(pfb_ctx *ctx, uint64_t *n_records, uint64_t *n_picked) {
    if (!ctx) return PFB_EINVAL;
    if (!ctx->computed) return fail(ctx, PFB_ESTATE, "no completed run");
    if (n_records) *n_records = ctx->nrec;
    if (n_picked) *n_picked = ctx->npick;
    return ctx->nrec ? PFB_OK : PFB_ENOSHARED;
}

int pfb_result_records(pfb_ctx *ctx, pfb_record *out, uint64_t cap) {
    if (!ctx || (!out && cap)) return PFB_EINVAL;
    if (!ctx->computed) return fail(ctx, PFB_ESTATE, "no completed run");
    if (cap < ctx->nrec) return fail(ctx, PFB_EINVAL, "record buffer too small");
    CK(cudaSetDevice(ctx->device));
    if (ctx->nrec) {
        CK(cudaMemcpyAsync(out, ctx->rec.p, ctx->nrec * sizeof(pfb_record), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return PFB_OK;
}

int pfb_result_positions(pfb_ctx *ctx, uint32_t genome_idx, int picked_only, pfb_pos *out, uint64_t cap) {
    if (!ctx || (!out && cap)) return PFB_EINVAL;
    if (!ctx->computed) return fail(ctx, PFB_ESTATE, "no completed run");
    if (genome_idx >= ctx->genomes.size()) return fail(ctx, PFB_EINVAL, "genome index out of range");
    CK(cudaSetDevice(ctx->device));
    const uint64_t need = picked_only ? ctx->npick : ctx->nrec;
    if (cap < need) return fail(ctx, PFB_EINVAL, "position buffer too small");
    if (!ctx->nrec) return PFB_OK;
    const pfb_pos *src = (const pfb_pos *)ctx->pos.p + (size_t)genome_idx * ctx->nrec;
    if (!picked_only) {
        CK(cudaMemcpyAsync(out, src, ctx->nrec * sizeof(pfb_pos), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return PFB_OK;
    }
    // picked only: copy all, compact on the host side of the ABI (the pick vector is tiny)
    std::vector<pfb_pos> all(ctx->nrec);
    std::vector<uint8_t> pk(ctx->nrec);
    CK(cudaMemcpyAsync(all.data(), src, ctx->nrec * sizeof(pfb_pos), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(pk.data(), ctx->pick.p, ctx->nrec, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint64_t n = 0;
    for (uint64_t r = 0; r < ctx->nrec; r++) if (pk[r]) out[n++] = all[r];
    return PFB_OK;
}

int pfb_timings(pfb_ctx *ctx, pfb_timing *out) {
    if (!ctx || !out) return PFB_EINVAL;
    *out = ctx->tm;
    return PFB_OK;
}

int pfb_set_debug_counts(pfb_ctx *ctx, int on) {
    if (!ctx) return PFB_EINVAL;
    ctx->dbg_counts = on ? 1 : 0;
    return PFB_OK;
}

int pfb_first_genome_allowed(pfb_ctx *ctx, uint64_t *out, uint32_t cap) {
    if (!ctx || !out) return PFB_EINVAL;
    if (ctx->stage_done < 0) return fail(ctx, PFB_ESTATE, "stage 0 has not run");
    if (cap < ctx->allowed.size()) return fail(ctx, PFB_EINVAL, "buffer too small");
    for (size_t i = 0; i < ctx->allowed.size(); i++) out[i] = ctx->allowed[i];
    return PFB_OK;
}

int pfb_oligo_tm(pfb_ctx *ctx, const char *seq, double *tm, double *gc_percent) {
    if (!ctx || !seq) return PFB_EINVAL;
    size_t n = strlen(seq);
    if (n < 2 || n > 32) return fail(ctx, PFB_EINVAL, "oligo length must be 2..32");
    uint64_t h = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t c;
        switch (seq[i]) { case 'A': c = 0; break; case 'T': c = 1; break; case 'C': c = 2; break; case 'G': c = 3; break;
            default: return fail(ctx, PFB_EINVAL, "non-ACGT base"); }
        h = (h << 2) | c;
    }
    CK(cudaSetDevice(ctx->device));
    // oligotm constants (same derivation as stage0)
    K k = ctx->k;
    const pfb_params &P = ctx->prm;
    double dv = P.dv_conc, dntp = P.dntp_conc;
    if (dv == 0) dntp = 0;
    if (dv < dntp) dv = dntp;
    k.ln_salt = std::log((P.mv_conc + 120.0 * std::sqrt(dv - dntp)) / 1000.0);
    k.ln_dna4 = std::log(P.dna_conc / 4000000000.0);
    k.ln_dna1 = std::log(P.dna_conc / 1000000000.0);
    k.dmso_term = P.dmso_conc * P.dmso_fact; k.formamide = P.formamide_conc;
    int rc;
    if ((rc = ensure(ctx, ctx->misc, 4096)) != PFB_OK) return rc;
    double *d = (double *)((uint8_t *)ctx->misc.p + 3072);
    k_oligo<<<1, 1, 0, ctx->stream>>>(k, h, (int)n, d);
    double hres[2];
    CK(cudaMemcpyAsync(hres, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (tm) *tm = hres[0];
    if (gc_percent) *gc_percent = hres[1];
    return PFB_OK;
}

}  // extern "C"
