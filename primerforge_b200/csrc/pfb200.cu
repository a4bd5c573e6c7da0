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
//   * genome 1's CANDIDATE windows (one end G/C, ACGT only, and -- with prefilter -- GC / repeat / Tm
//     in range) are entered into a 2-bit-per-bin table (seen once / seen again) with L2-resident
//     atomicOr (k_cand1).  A candidate that is alone in its bin owns the bin, so (k, bin) is a
//     perfect hash of the key set: a bitmap over bins + popcount ranks gives a dense slot id
//     (k_candbits, k_scan_p*) -- this replaces the string sets and the Aho-Corasick
//     automaton;
//   * every genome, the first included, is then ONE pass (k_scan_filtered2 / k_scan_filtered / k_scan_plain):
//     window -> bin -> key-table probe; a window landing in a key's bin adds (1 << 32 | start) to the
//     (genome, key) row with a fire-and-forget RED; windows touching a '~' junction or a non-ACGT byte
//     pollute only one strand's table (two flag bits; k_junction).  After the pass "count == 1" is khmer's
//     get() == 1 and the single occupant is read back and compared with the key (k_alive1 / k_aliveN): a
//     key is shared iff every genome has exactly one occurrence and a clean bin in the strand table the
//     reference would have read.  The probe goes through a per-k Bloom-style filter held in shared memory
//     (179 KB per CTA, bulk-copied per length: pfb_bulk.cuh) so that only ~1 in 5 windows reaches L2;
//   * genome 1 is scanned first; its survivors get ids in the reference's dict order (genome-1 end
//     position ascending, longest first) and everything downstream -- rows of the other genomes, verdicts,
//     records, coordinates -- is indexed by id and streams through memory (k_alive1,
//     k_build_at, k_aliveN, k_id_records); filters are evaluated in fp64 with the reference's operation
//     order, the per-position "first k-mer that passes" pick looks only at the few records that can share
//     a position (k_pick); the end of a run is one launch (k_finish);
//   * no step waits for the host: capacities are known to the host, counts stay on the device; results stream to
//     pinned host memory while the run goes on;
//   * more than one GPU: genomes are dealt out, genome 1's own pipeline is shared out by position, and what the ranks
//     exchange (candidate lists, packed rows, alive / pick bytes) is read from the peers' memory over NVLink, or
//     goes through NCCL;
//   * FASTA files can be handed over as they are: records and sequence lines are found on the device
//     (pfb_fasta.cuh).
//
// No tensor cores: nothing here is a contraction.

#include "../../include/pfb200.h"
#include "pfb_fasta.cuh"
#include "pfb_bulk.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// row word layout (one uint64 per genome and slot).  Every window of the genome that lands in the slot's bin
// ADDs (1 << 32) | start: after the scan "count == 1" is khmer's get() == 1 for that bin and, only then, the
// low word is the start of the single occupant, which k_alive compares with the slot's key.  A fire-and-forget
// RED: the scan never waits for an atomic's return value.  Strand-specific polluters OR the two top bits.
constexpr uint64_t ROW_POS_MASK = 0x03FFFFFFull;   // leftmost start, genome-local padded coordinate
constexpr uint64_t ROW_ONE = 1ull << 32;           // occupancy count lives in bits 32..59
constexpr uint64_t ROW_CNT_MASK = 0x0FFFFFFFull;
constexpr uint64_t ROW_JF = 1ull << 62;            // forward-string-only polluter ('~' or non-ACGT window)
constexpr uint64_t ROW_JR = 1ull << 63;            // reverse-string-only polluter
constexpr uint32_t MAX_GENOME_WORDS = 1u << 21; // 2^26 padded bases
constexpr uint32_t ID_NONE = 0xFFFFFFFFu;

constexpr uint32_t FILTER_BYTES = 179u * 1024u; // shared-memory filter of the filtered scans
constexpr uint32_t FILTER_BITS = FILTER_BYTES * 8u;
constexpr uint32_t FILTER_WORDS = FILTER_BYTES / 4u;
constexpr uint32_t QCAP = 64;                   // per-warp queue of filter-positive windows (16 B items)
constexpr uint32_t SCANF_SMEM = FILTER_BYTES + 32u * QCAP * 16u;
// k_scan_filtered2: per warp a stack of 96 16-byte items (two windows per lane may arrive between two "batch ready" tests)
constexpr uint32_t Q2_ITEMS = 96;
constexpr uint32_t SCAN2_SMEM = FILTER_BYTES + 32u * Q2_ITEMS * 16u;
static_assert(SCAN2_SMEM <= 227u * 1024u, "shared memory of k_scan_filtered2");
constexpr int SCAN_TILE = 2048;                 // elements per block of the device-wide scans

struct K {  // kernel arguments (by value, __grid_constant__)
    int kmin, kmax, nk;
    int kiSplit;              // the filtered scan: lengths [0, kiSplit) by k_scan_filtered2, [kiSplit, nk) by k_scan_filtered
    uint32_t P, PW;           // table prime, words per 1-bit bitmap (the 2-bit table has 2 * PW words)
    uint64_t barrett;         // floor(2^64 / P)
    uint32_t fscale;          // floor(FILTER_WORDS * 2^32 / P): bin -> filter word (the bit is bin & 31)
    int max_repeat, prefilter;
    double min_tm, max_tm;
    double ln_salt, ln_dna4, ln_dna1, dmso_term;
    uint64_t gc_ok[33];       // per k: bit c set <=> gc count c is inside [minGc, maxGc]
    const double *fterm;      // [33][33] formamide term per (k, gc count)
    const float4 *tmb;        // [33][33] Tm window as two lines in (ds, dh): dh >= x*ds + y and dh <= z*ds + w
    // arena
    const uint8_t *raw;
    uint64_t *seq2;
    uint32_t *nmask;
    uint32_t *wordContig;
    const uint32_t *cWordStart, *cLen, *cGenome, *cVirt;
    const uint64_t *cRawStart;
    const uint32_t *gWordStart, *gFirstContig, *gNumContigs, *gVirtLen;
    uint32_t nContigs, nGenomes, totalWords;
    // genome-1 candidate table + key table
    uint32_t *sk;             // [nk][2*PW] 2 bits per bin: seen once (low half), seen again (high half)
    uint2 *br;                // [nk][PW] .x = key bitmap word, .y = exclusive popcount prefix (slot id base)
    uint32_t *filter;         // [nk][FILTER_WORDS]
    uint32_t *emitMask;       // per genome-1 end position: bit per k = surviving key ending there
    uint32_t *emitOff;        // exclusive popcount prefix of emitMask = record index base
    // rows: genome 1's row is indexed by slot (bin order); after its scan the slots that are still unique
    // there get an ID in the reference's dict order (genome-1 end position, longest first), and the rows of
    // every other genome, the records and the coordinates are indexed by id: everything after the scan
    // streams through memory in order
    unsigned long long *rows0; // [nslots]
    unsigned long long *rows;  // [nGenomes - 1][nids]
    uint32_t *remap;           // [nslots] slot -> id, ID_NONE = not unique in genome 1
    uint2 *at;                 // [nk][PW16] key table of the other genomes: per 16 bins (bitmap16 | id0:24 | id1:24)
    uint32_t *ovf;             // [n][16] ids of the words that do not fit that format
    uint32_t *ovfCount;
    uint32_t PW16, ovfCap;
    uint32_t atWide;           // ids from here on do not fit the inline format (AT_SLOW; lowered by tests)
    int first_pass;            // 1 while genome 1 itself is scanned (hits go to rows0[slot]); 2: the same for one rank's
                               // share of genome 1 (its polluter windows are handled by k_scan_plain<true> on every rank)
    // The host never waits for a count in the middle of a run: buffers have CAPACITIES the host knows (capSlots,
    // capIds = the pitch of every per-id row) and the kernels read the actual counts from device memory (cnt).  A
    // count that exceeds its capacity raises a flag in cnt[C_FLAGS]; the host sees it at the end, grows and reruns.
    uint32_t *cnt;             // device counters, C_* below (mirrored in host-mapped memory by the kernels that total them)
    uint32_t capSlots, capIds;
    uint64_t *key1;            // [capIds] the k-mer as it reads on genome 1's (+) strand
    uint32_t *idMeta;          // [capIds] genome-1 start | (k - kmin) << 26
    uint8_t *alive;            // [capIds] 1 while unique in every genome checked so far
    // output, indexed by id (= the reference's dict order)
    double *idTm;              // [capIds]
    uint16_t *idInfo;          // [capIds] (k - 1) | gc_count << 5 | flags << 11
    uint32_t *pos32;           // [nGenomes][capIds] genome-local padded coordinate of the leftmost base | strand << 31
    uint8_t *pick;             // [capIds] 0 / 1
    // multi-GPU: genome 1's own pipeline is shared out by position
    uint32_t *candList;        // [nRanks][capCand + 4]: segment r = {count, 0, 0, 0, (bin | ki << 27) ...} of rank r
    uint32_t capCand, nRanks, rank;
    uint32_t *rows0p;          // [capSlots] rows0 packed to 32 bits for the SUM all-reduce
    // multi-GPU over peer memory (NVLink): what the ranks exchange -- candidate segment, packed rows, the reduced
    // slice of the packed rows, alive bytes, pick bytes -- lives in ONE arena per rank that every other rank has
    // mapped (cudaIpc* between processes, peer access inside one); consumers read the peers' arenas directly, a
    // flag per (phase, rank) written into every peer's arena says when a producer is done
    uint8_t *peer[16];         // the arenas, peer[rank] = this rank's own
    uint32_t xRows0p, xRows0s, xAlive, xPick, xCand;   // byte offsets of the parts within an arena (flags at 0)
    uint32_t xEpoch;           // number of this run on the communicator: the value the flags must reach
};

enum { XP_CANDS = 0, XP_ROWS_A = 1, XP_ROWS_B = 2, XP_ALIVE = 3, XP_PICK = 4, XP_DONE = 5, XP_PHASES = 6 };
constexpr uint32_t X_FLAGS_BYTES = 4096;       // [XP_PHASES][16] uint32 (+ slack) at the start of an arena

enum { C_NSLOTS = 0, C_NIDS = 1, C_NALIVE = 2, C_NPICK = 3, C_OVF = 4, C_FLAGS = 5, C_MAXCAND = 6, C_TICKET = 7 };
enum { OVF_SLOTS = 1u, OVF_IDS = 2u, OVF_LISTS = 4u, OVF_CAND = 8u, OVF_PEER_TIMEOUT = 16u };

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

__device__ __forceinline__ int gc_count(uint64_t h) { return __popcll(h & 0xAAAAAAAAAAAAAAAAull); }

// Tm from the nearest-neighbour sums of a k-mer (acc = (sum S << 16) | sum H over its k-1 pairs).
// Operation order mirrors primer3's oligotm() as restated in oracle/pf_oracle.c so that the doubles
// agree bit for bit (explicit _rn intrinsics: no FMA contraction).
__device__ __forceinline__ double tm_from_sums(uint32_t acc, uint64_t h, int k, int gc, const K &p) {
    uint64_t rc = rc64(h) >> (64 - 2 * k);
    bool sym = ((k & 1) == 0) && rc == h;
    int dh = (int)(acc & 0xFFFF), ds = (int)(acc >> 16);
    if (sym) ds += 14;
    int first = (int)((h >> (2 * k - 2)) & 3), last = (int)(h & 3);
    if (first < 2) { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    if (last < 2)  { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    double delta_h = __dmul_rn((double)dh, -100.0);
    double delta_s = __dmul_rn((double)ds, -0.1);
    delta_s = __dadd_rn(delta_s, __dmul_rn(__dmul_rn(0.368, (double)(k - 1)), p.ln_salt));
    double ln_dna = sym ? p.ln_dna1 : p.ln_dna4;
    double tm = __dsub_rn(__ddiv_rn(delta_h, __dadd_rn(delta_s, __dmul_rn(1.987, ln_dna))), 273.15);
    tm = __dsub_rn(tm, p.dmso_term);
    tm = __dadd_rn(tm, p.fterm[k * 33 + gc]);
    return tm;
}

// minTm <= Tm <= maxTm (:305-307).  For a fixed length and GC count, Tm >= t is the half plane
// dh >= a * ds + b of the integer nearest-neighbour sums (the denominator of the formula is negative), so
// the window test is two fused multiply-adds against a per-(k, gc) table.  dh is an integer: unless it is
// within 0.004 of a bound the fp32 evaluation (error < 5e-4: |ds| < 9000, line slopes ~0.33, offsets ~180)
// decides; otherwise -- and for self-complementary oligos, which use another constant -- the exact fp64
// formula does, so the decision is always the one the fp64 formula makes.
__device__ __forceinline__ bool tm_in_range(uint32_t acc, uint64_t h, int k, int gc, const K &p, bool self_compl) {
    const uint32_t first = (uint32_t)(h >> (2 * k - 2)) & 3u, last = (uint32_t)h & 3u;
    const int dh = (int)(acc & 0xFFFF) + (first < 2 ? -23 : -1) + (last < 2 ? -23 : -1);
    const int ds = (int)(acc >> 16) + (first < 2 ? -41 : 28) + (last < 2 ? -41 : 28);
    if (!self_compl) {
        const float4 b = __ldg(&p.tmb[k * 33 + gc]);
        const float x = (float)ds, d = (float)dh;
        const float lo = fmaf(b.x, x, b.y), hi = fmaf(b.z, x, b.w);
        if (d < lo - 0.004f || d > hi + 0.004f) return false;
        if (d > lo + 0.004f && d < hi - 0.004f) return true;
    }
    double tm = tm_from_sums(acc, h, k, gc, p);
    return tm >= p.min_tm && tm <= p.max_tm;
}

// nn = the 16-entry table in shared memory (bank = index: divergent lookups never conflict)
__device__ __forceinline__ double oligo_tm(uint64_t h, int k, int gc, const K &p, const uint32_t *nn) {
    uint32_t acc = 0;
#pragma unroll 4
    for (int i = 0; i < k - 1; i++) acc += nn[(h >> (2 * i)) & 15];
    return tm_from_sums(acc, h, k, gc, p);
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

__device__ __forceinline__ uint32_t filter_flags(uint64_t h, int k, const K &p, const uint32_t *nn, double *tm_out) {
    uint32_t f = 0;
    int gc = gc_count(h);
    if ((p.gc_ok[k] >> gc) & 1ull) f |= PFB_F_GC_OK;                      // getCandidateKmers.py:301-303
    if (!has_long_repeat(h, k, p.max_repeat)) f |= PFB_F_REP_OK;          // :309-320
    double tm = oligo_tm(h, k, gc, p, nn);
    if (tm >= p.min_tm && tm <= p.max_tm) f |= PFB_F_TM_OK;               // :305-307
    *tm_out = tm;
    return f;
}

// key table probe: one 8-byte load gives the bitmap word and the slot id base
__device__ __forceinline__ uint2 probe(const K &p, int ki, uint32_t bin) {
    return __ldg(&p.br[(size_t)ki * p.PW + (bin >> 5)]);
}
__device__ __forceinline__ bool probe_hit(uint2 e, uint32_t bin) { return (e.x >> (bin & 31)) & 1u; }
__device__ __forceinline__ uint32_t slot_of(uint2 e, uint32_t bin) { return e.y + __popc(e.x & ((1u << (bin & 31)) - 1u)); }

// ------------------------------------------------------------------------------------------
// k_encode: ASCII -> 2 bits / base (+ invalid mask); one thread per 32-base word.
// Replaces str(rec.seq) / reverse_complement / '~'.join (getCandidateKmers.py:93-100).
__global__ void __launch_bounds__(256) k_encode(const __grid_constant__ K p, uint32_t w0, uint32_t w1) {
    uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= w1) return;
    // owning contig: last contig whose first word is <= w.  The 32 words of a warp nearly always lie in one
    // contig: the search runs once for the warp's first word (the same loads in every lane) and again per
    // lane only when the next contig starts inside the warp's span
    const uint32_t w_first = w - (threadIdx.x & 31u);
    uint32_t lo = 0, hi = p.nContigs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.cWordStart[mid] <= w_first) lo = mid; else hi = mid;
    }
    if (lo + 1 < p.nContigs && p.cWordStart[lo + 1] <= w) {
        hi = p.nContigs;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.cWordStart[mid] <= w) lo = mid; else hi = mid;
        }
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
        inval |= 0xFFFFFFFFu >> nvalid;
        packed |= ~0ull >> (2 * nvalid);
    }
    p.seq2[w] = packed;
    p.nmask[w] = inval;
    p.wordContig[w] = c;
}

// ------------------------------------------------------------------------------------------
// bookkeeping of one probe hit (any window of genome g whose bin holds a key)
// Key lookup.  Genome 1 is scanned against `br` (bitmap word + popcount rank -> slot, rows0[slot]).  The other
// genomes are scanned against `at`, built after genome 1's scan from the keys that are still alive: one
// 8-byte word per 16 bins holding the bitmap of the alive keys and, inline, the ids of the first two of them
// (ids follow the reference's dict order, not the bin order, so they cannot be a popcount rank).  With keys in
// ~2 % of the bins more than two per 16 bins is rare (< 1 % of the words); those words carry AT_SLOW and the
// index of a 16-entry id list.  One load, no dependent second load on the hot path.
constexpr uint32_t AT_SLOW = 0xFFFFFFu;           // id0 field: "the ids are in ovf[e.y >> 8]"

__device__ __forceinline__ uint2 probe2(const K &p, int ki, uint32_t bin) {
    return __ldg(&p.at[(size_t)ki * p.PW16 + (bin >> 4)]);
}

// id of the alive key in bin (b16 = bin & 15) given its probe word, ID_NONE when the bin holds none
__device__ __forceinline__ uint32_t id_of(const K &p, uint2 e, uint32_t b16) {
    if (!((e.x >> b16) & 1u)) return ID_NONE;
    const uint32_t id0 = (e.x >> 16) | ((e.y & 0xFFu) << 16);
    if (id0 == AT_SLOW) return __ldg(&p.ovf[(size_t)(e.y >> 8) * 16u + b16]);
    return (e.x & ((1u << b16) - 1u)) ? (e.y >> 8) : id0;
}

// the row a window of genome g in `bin` counts into (nullptr: no key there)
__device__ __forceinline__ unsigned long long *row_of_bin(const K &p, int ki, uint32_t g, uint32_t bin) {
    if (p.first_pass) {
        const uint2 e = probe(p, ki, bin);
        if (!probe_hit(e, bin)) return nullptr;
        const uint32_t slot = slot_of(e, bin);
        return slot < p.capSlots ? &p.rows0[slot] : nullptr;      // (beyond the capacity: flagged by k_alive1, the run is repeated)
    }
    const uint32_t id = id_of(p, probe2(p, ki, bin), bin & 15u);
    return id == ID_NONE ? nullptr : &p.rows[(size_t)(g - 1) * p.capIds + id];
}

__device__ __forceinline__ void pollute(const K &p, int ki, uint32_t g, uint64_t canon, bool rev_table) {
    unsigned long long *row = row_of_bin(p, ki, g, mod_p(canon, p));
    if (row) atomicOr(row, rev_table ? ROW_JR : ROW_JF);
}

// window with >= 1 non-ACGT byte: a strand-specific polluter.  The forward string hashes the byte as
// (3, comp 2); the reverse-complement string keeps the byte, so there it is again (3, comp 2).
__device__ __noinline__ void pollute_invalid_window(const K &p, int ki, int k, uint32_t g, uint64_t h, uint64_t r, uint32_t wn) {
    uint64_t canon_f = h < r ? h : r;
    uint64_t h2 = r, r2 = h;
    for (int t = 0; t < k; t++)
        if ((wn >> t) & 1u) { h2 |= 1ull << (2 * (k - 1 - t)); r2 &= ~(1ull << (2 * t)); }
    uint64_t canon_r = h2 < r2 ? h2 : r2;
    pollute(p, ki, g, canon_f, false);
    pollute(p, ki, g, canon_r, true);
}

// ------------------------------------------------------------------------------------------
// genome-1 kernels: one thread per window END position (coalesced, 32x the parallelism of a
// per-word walk; genome 1 is a single genome, so parallelism matters more than rolling reuse)
struct PosState {
    uint64_t hf, hr;   // last 32 bases ending here: forward word (newest base lowest), rc word (top aligned)
    uint64_t nmw;      // invalid-base bits, bit t = base of age t
    uint32_t g, gpos;  // genome, genome-local padded coordinate of this end base
    int n;             // bases of the contig up to and including this one
};

__device__ __forceinline__ bool load_pos(const K &p, uint32_t b, PosState &s) {
    uint32_t w = b >> 5, j = b & 31;
    uint32_t c = p.wordContig[w];
    uint32_t i = w - p.cWordStart[c];
    uint32_t q = 32u * i + j;
    if (q >= p.cLen[c]) return false;
    uint64_t cur = p.seq2[w], prev = i ? p.seq2[w - 1] : 0ull;
    s.hf = j == 31 ? cur : ((prev << (2 * (j + 1))) | (cur >> (62 - 2 * j)));
    s.hr = rc64(s.hf);
    uint64_t nm = ((uint64_t)(i ? p.nmask[w - 1] : 0u) << 32) | p.nmask[w];
    s.nmw = nm >> (31 - j);
    s.g = p.cGenome[c];
    s.gpos = b - 32u * p.gWordStart[s.g];
    s.n = (int)q + 1;
    return true;
}

// k_cand1: candidate windows of genome 1 -> 2-bit table.  Candidate = a k-mer the reference could
// keep: no '~' / non-ACGT (:40-43), one end G/C (:29-33) and, with prefilter, GC / repeat / Tm in
// range (:301-320; legal because the per-position pick takes the FIRST k-mer that passes, so k-mers
// that fail can be dropped up front).  The sketch's "count == 1" is decided later by the scan, which
// sees every window of genome 1, candidates or not.
//
// A CTA covers CAND_POS consecutive end positions (one per thread) plus a 32-position halo warp.
//   1. every thread looks up the nearest-neighbour value of the base pair ending at its position; a CTA-wide
//      prefix sum P of those values turns the sum over ANY window into one subtraction, P[b] - P[b - k + 1]
//      (both fields of the packed (S << 16 | H) value are differences of < 2^14, so the 32-bit wrap-around of
//      the prefix is harmless);
//   2. the cheap tests (ends, non-ACGT, GC window, homopolymers) run branch-free for every k;
//   3. the (position, k) pairs that survive (~38 %) are compacted per warp, so that the Tm test, the
//      canonical hash and the table update run with full warps instead of a few live lanes per k.
constexpr int CAND_POS = 256;

// Multi-GPU (LIST): a rank works on its share of genome 1's CTAs and, instead of updating the table, appends
// (bin | ki << 27) to its segment of candList; the segments are all-gathered and every rank applies all of them
// (k_apply_cands), so the tables end up identical everywhere -- the order of the updates does not matter.
__device__ __forceinline__ void cand_table_update(const K &p, int ki, uint32_t bin) {
    uint32_t *cell = &p.sk[(size_t)ki * 2 * p.PW + (bin >> 4)];
    const uint32_t bbit = bin & 15u;
    const uint32_t old = atomicOr(cell, 1u << bbit);
    if ((old >> bbit) & 1u) atomicOr(cell, 1u << (16 + bbit));
}

template <bool LIST>
__global__ void __launch_bounds__(CAND_POS + 32) k_cand1(const __grid_constant__ K p, uint32_t nbases, uint32_t cta0) {
    __shared__ uint32_t s_nn[16];
    __shared__ uint32_t s_P[CAND_POS + 32];
    __shared__ uint32_t s_wsum[(CAND_POS + 32) / 32];
    __shared__ uint64_t s_km2[32];
    __shared__ uint32_t s_gclo[32], s_gcspan[32];     // GC counts inside [minGc, maxGc] form an interval per k
    __shared__ uint16_t s_item[(CAND_POS + 32) / 32][32 * 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) s_nn[tid] = c_nn[tid];
    const int R = p.max_repeat;
    if (tid >= 32 && tid < 32 + (uint32_t)p.nk) {
        const int ki = tid - 32, k = p.kmin + ki;
        s_km2[ki] = kmask2(k);
        const uint64_t ok = p.gc_ok[k];
        s_gclo[ki] = ok ? (uint32_t)(__ffsll((long long)ok) - 1) : 64u;
        s_gcspan[ki] = ok ? (uint32_t)(63 - __clzll((long long)ok)) - s_gclo[ki] : 0u;
    }
    __syncthreads();
    // thread t <-> end position b = first position of the CTA - 32 + t (warp 0 is the halo)
    const long long bb = (long long)(blockIdx.x + cta0) * CAND_POS - 32 + tid;
    PosState s;
    s.hf = 0; s.hr = 0; s.nmw = ~0ull; s.n = 0; s.g = 0; s.gpos = 0;
    const bool have = bb >= 0 && bb < (long long)nbases && load_pos(p, (uint32_t)bb, s);
    const uint64_t hf = s.hf;
    // 1. prefix sums of the pair values
    uint32_t incl = have ? s_nn[(uint32_t)hf & 15u] : 0u;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t i = 0; i < warp; i++) woff += s_wsum[i];
    s_P[tid] = incl + woff;
    __syncthreads();
    if (warp == 0) return;                      // the halo only contributes its pair values
    // 2. the cheap tests.  Three of them only bound k from above and fold into ONE limit: the contig start (k <= n,
    // :40-43), the newest non-ACGT base (age d: k <= d) and, with prefilter, the newest homopolymer run of
    // max_repeat bases (starting at age i: k <= i + R - 1, :309-320).  What is left per k is one bit of the word
    // -- is the FIRST base of the window G/C? -- which serves both isOneEndGc (:29-33) and a running GC count
    // that is compared with the per-k interval (:301-303).  Survivors are compacted per warp right away.
    const bool pre = p.prefilter != 0;
    int khi = have ? min(p.kmax, s.n) : 0;
    const uint32_t nm32 = (uint32_t)s.nmw;                           // bit t = the base of age t is not A C G T
    if (nm32) khi = min(khi, __ffs((int)nm32) - 1);
    if (pre) {
        if (R <= 1) {
            khi = 0;
        } else {
            const uint64_t e = hf ^ (hf >> 2);
            const uint64_t z = ~(e | (e >> 1)) & 0x5555555555555555ull;     // bit 2i: ages i and i + 1 hold the same base
            uint64_t runs = z;
            for (int t = 1; t < R - 1; t++) runs &= z >> (2 * t);           // bit 2i: a run of R bases starts at age i
            if (runs) khi = min(khi, ((__ffsll((long long)runs) - 1) >> 1) + R - 1);
        }
    }
    const uint32_t last_gc = (uint32_t)(hf >> 1) & 1u;
    const uint64_t ends = hf >> (2 * p.kmin - 1);                           // bit 2 ki: first base of window kmin + ki is G/C
    uint32_t gc = p.kmin > 1 ? (uint32_t)__popcll(hf & 0xAAAAAAAAAAAAAAAAull & kmask2(p.kmin - 1)) : 0u;
    uint16_t *items = s_item[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t n_items = 0;
    for (int ki = 0; ki < p.nk; ki++) {
        const int k = p.kmin + ki;
        const uint32_t b = (uint32_t)(ends >> (2 * ki)) & 1u;
        gc += b;
        bool ok = (k <= khi) & ((last_gc | b) != 0u);
        if (pre) ok &= (gc - s_gclo[ki]) <= s_gcspan[ki];
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
        if (ok) items[n_items + __popc(m & lt_mask)] = (uint16_t)(lane | (ki << 5));
        n_items += __popc(m);
    }
    __syncwarp();
    const uint32_t hf_lo = (uint32_t)hf, hf_hi = (uint32_t)(hf >> 32);
    uint32_t *seg = !LIST ? nullptr : p.xCand ? (uint32_t *)(p.peer[p.rank] + p.xCand) : p.candList + (size_t)p.rank * (p.capCand + 4u);
    for (uint32_t base = 0; base < n_items; base += 32) {
        const bool live = base + lane < n_items;
        const uint32_t it = live ? items[base + lane] : 0u;
        const uint32_t src = it & 31u;
        const int ki = (int)(it >> 5), k = p.kmin + ki;
        const uint64_t w = ((uint64_t)__shfl_sync(0xFFFFFFFFu, hf_hi, src) << 32) | __shfl_sync(0xFFFFFFFFu, hf_lo, src);
        bool ok = live;
        uint32_t bin = 0;
        if (ok) {
            const uint64_t h = w & s_km2[ki];
            const uint64_t r = rc64(h) >> (64 - 2 * k);
            if (pre) {
                const uint32_t pi = warp * 32 + src;
                const uint32_t acc = s_P[pi] - s_P[pi - (uint32_t)(k - 1)];
                ok = tm_in_range(acc, h, k, __popcll(h & 0xAAAAAAAAAAAAAAAAull), p, r == h);
            }
            if (ok) bin = mod_p(h < r ? h : r, p);
        }
        if (LIST) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
            if (m) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(seg, (uint32_t)__popc(m));
                at = __shfl_sync(0xFFFFFFFFu, at, 0) + __popc(m & lt_mask);
                if (ok && at < p.capCand) seg[4u + at] = bin | ((uint32_t)ki << 27);   // (a full segment: flagged by k_apply_cands)
            }
        } else if (ok) {
            cand_table_update(p, ki, bin);
        }
    }
}

// every rank's candidates -> this rank's table
template <bool PEER>
__global__ void __launch_bounds__(256) k_apply_cands(const __grid_constant__ K p) {
    // one entry per thread (blockIdx.y = the rank whose segment it is): every table update of the launch is in
    // flight at once -- the updates return a value, a thread that made several in a row would wait for each
    const uint32_t r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    // PEER: rank r's segment is read where it was written, in r's arena, over NVLink (no all-gather at all)
    const uint32_t *seg = PEER ? (const uint32_t *)(p.peer[r] + p.xCand) : p.candList + (size_t)r * (p.capCand + 4u);
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = __ldcg(seg);               // (one read of the count per CTA, not one per thread)
    __syncthreads();
    uint32_t n = s_n;
    if (n > p.capCand) { if (i == 0) { atomicOr(&p.cnt[C_FLAGS], OVF_CAND); atomicMax(&p.cnt[C_MAXCAND], n); } n = p.capCand; }
    if (i >= n) return;
    const uint32_t e = __ldcg(seg + 4u + i);
    cand_table_update(p, (int)(e >> 27), e & 0x07FFFFFFu);
}

// ------------------------------------------------------------------------------------------
// k_scan_plain: every window of every genome against the key table, probing the bitmap in L2
// directly.  Used when the key set is too dense for the filter (prefilter = 0) and on tiny inputs.
// One thread per 32-base word = 32 end positions x nk lengths with rolling words.
// Implements sketch membership (:54-55), the strand algebra (:58-63), the per-k intersection
// (:127-142) and the relocation (:160-187) in one probe.
// POLLUTERS: only the windows with a non-ACGT byte (the strand-specific polluters) -- run by every rank over all
// of genome 1 after a shared-out first pass (first_pass == 2), whose own scans leave them alone.
template <bool POLLUTERS>
__global__ void __launch_bounds__(256) k_scan_plain(const __grid_constant__ K p, uint32_t w0, uint32_t w1) {
    uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= w1) return;
    uint32_t c = p.wordContig[w];
    uint32_t cws = p.cWordStart[c], clen = p.cLen[c], g = p.cGenome[c];
    uint32_t i = w - cws;
    uint64_t nm = ((uint64_t)(i ? p.nmask[w - 1] : 0u) << 32) | p.nmask[w];
    if (POLLUTERS && nm == 0) return;
    uint64_t cur = p.seq2[w], prev = i ? p.seq2[w - 1] : 0ull;
    uint64_t hf = prev, hr = rc64(prev);
    uint32_t qbase = 32u * i;
    int nj = (int)min(32u, clen - qbase);
    uint32_t gpos0 = 32u * (w - p.gWordStart[g]);
    // (probing four lengths at once with the 32-bit modular reduction was tried: slower, 0.280 vs 0.258 ms for
    // one 5 Mbp genome -- the kernel lives on thread-level parallelism, the extra registers cost occupancy)
    for (int j = 0; j < nj; j++) {
        uint32_t code = (uint32_t)(cur >> (62 - 2 * j)) & 3u;
        hf = (hf << 2) | code;
        hr = (hr >> 2) | ((uint64_t)(code ^ 1u) << 62);
        int khi = min(p.kmax, (int)qbase + j + 1);
        const uint64_t nmw = nm >> (31 - j);
        for (int k = p.kmin; k <= khi; k++) {
            int ki = k - p.kmin;
            uint64_t h = hf & kmask2(k), r = hr >> (64 - 2 * k);
            uint32_t wn = (uint32_t)(nmw & kmask1(k));
            if (wn) { if (POLLUTERS || p.first_pass != 2) pollute_invalid_window(p, ki, k, g, h, r, wn); continue; }
            if (POLLUTERS) continue;
            uint64_t canon = h < r ? h : r;
            uint32_t bin = mod_p(canon, p);
            unsigned long long *row = row_of_bin(p, ki, g, bin);
            if (row) atomicAdd(row, ROW_ONE | (gpos0 + j - k + 1));
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_scan_filtered: the hot kernel.  One persistent 1024-thread CTA per SM.  For each k the CTA
// copies that k's 192 KB filter (1 bit per ~6.4 bins, set when a bin of the range holds a key) into
// shared memory and walks its share of the arena with rolling 2-bit words.  A window whose filter
// bit is clear is done after one shared-memory load (~4 of 5 windows).  The others are NOT probed
// in place: they are appended to a per-warp queue in shared memory, and whenever 32 are waiting the
// whole warp drains them together: the L2 probe (bitmap + rank word) is paid once per 32 windows with
// every lane busy, and a hit ends in one fire-and-forget RED on the genome's row.
// shared-memory accesses of the hot loop through 32-bit shared-window addresses (no generic-address
// conversion per access)
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));   // the filter is read-only during a k pass
    return v;
}
// predicated store: no divergent branch (and so no convergence barrier) around the queue append
__device__ __forceinline__ void sts128_if(uint32_t pred, uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.v4.u32 [%1], {%2, %3, %4, %5}; }"
                 ::"r"(pred), "r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// Drain = probe of 32 queued windows, software-pipelined: the probe loads of one batch are ISSUED when
// the batch fills up and CONSUMED when the next batch fills up (or at the end of the k pass), so their L2
// latency overlaps ~4 loop iterations of hashing instead of stalling the warp.
struct Pending {
    uint2 e;          // the probe word in flight: (bitmap word, rank)
    uint32_t meta;    // start | (bin & 31) << 26
    uint32_t g;
};

__device__ __forceinline__ void drain_issue(const K &p, int ki, uint32_t q_addr, uint32_t head, uint32_t lane, Pending &pd) {
    uint4 it = lds128(q_addr + (((head + lane) & (QCAP - 1)) << 4));   // x = bin, y = o, z = genome, w = start at o = 0
    pd.e = p.first_pass ? probe(p, ki, it.x) : probe2(p, ki, it.x);
    pd.meta = ((it.w - it.y) & (uint32_t)ROW_POS_MASK) | ((it.x & 31u) << 26);
    pd.g = it.z;
}

__device__ __forceinline__ void drain_consume(const K &p, const Pending &pd) {
    const uint32_t b = pd.meta >> 26;          // bin & 31: all the two tables need of the bin besides the probe word
    const unsigned long long add = ROW_ONE | (pd.meta & (uint32_t)ROW_POS_MASK);
    if (p.first_pass) {
        const uint32_t slot = pd.e.y + __popc(pd.e.x & ((1u << b) - 1u));
        if (((pd.e.x >> b) & 1u) && slot < p.capSlots) atomicAdd(&p.rows0[slot], add);
    } else {
        const uint32_t id = id_of(p, pd.e, b & 15u);
        if (id != ID_NONE) atomicAdd(&p.rows[(size_t)(pd.g - 1) * p.capIds + id], add);
    }
}

// 32-bit-word arithmetic of the hot loop.  A thread owns one packed word (32 window end positions) and
// its predecessor: 64 bases = the 128-bit stream prev : cur.  The window of k bases ending at base j of
// cur is bits [2o, 2o + 2k) of that stream (o = 31 - j) and its reverse complement is bits
// [128 - 2k - 2o, 128 - 2o) of the stream rc(cur) : rc(prev): each is ONE funnel shift per 32-bit
// half -- nothing is rolled base by base, every window is independent, and the shift amounts are the
// same for all lanes (uniform registers).  The loop is bound by the INT32 ALU pipe (profiles/
// r01_scan_ablation.md), so the code below is written to minimise ALU-pipe instructions: multiply-add
// forms for the modular reduction and the filter address (FMA pipe), no per-window validity work on
// words that are entirely inside a contig and free of non-ACGT bytes (CLEAN).
template <bool FASTMOD, bool KBIG>
__device__ __forceinline__ uint32_t canon_bin(const K &p, uint32_t h_lo, uint32_t h_hi, uint32_t r_lo, uint32_t r_hi,
                                              uint32_t m32, uint32_t a32, uint32_t neg_p) {
    uint32_t c_lo, c_hi;
    if (KBIG) {
        bool lt = ((((uint64_t)h_hi) << 32) | h_lo) < ((((uint64_t)r_hi) << 32) | r_lo);   // uniqify_rc: min(h, r)
        c_lo = lt ? h_lo : r_lo; c_hi = lt ? h_hi : r_hi;
    } else {
        c_lo = min(h_lo, r_lo); c_hi = 0;
    }
    if (FASTMOD) {
        // canon < 2^40 and P < 2^24: (c_hi * (2^32 mod P) + c_lo mod P) mod P in 32-bit multiply-adds
        uint32_t r1 = __umulhi(c_lo, m32) * neg_p + c_lo;                // c_lo - q P < 2P
        uint32_t s1 = KBIG ? c_hi * a32 + r1 : r1;                       // < 257 P < 2^32
        uint32_t r2 = KBIG ? __umulhi(s1, m32) * neg_p + s1 : s1;        // < 2P
        return min(r2, r2 + neg_p);                                      // unsigned: picks the reduced one
    }
    return mod_p(((uint64_t)c_hi << 32) | c_lo, p);
}

struct ScanWarp {            // per-warp state of k_scan_filtered
    uint32_t f_addr, q_addr; // shared-window addresses of the filter and of this warp's queue
    uint32_t qhead, qn;      // warp-uniform ring state
    bool pend;               // a batch of 32 probes is in flight
    Pending pd;
    uint32_t lane, lt_mask;
};

template <bool FASTMOD, bool KBIG, bool CLEAN>
__device__ __forceinline__ void scan_word(const K &p, int ki, int k, ScanWarp &sw,
                                          uint64_t prev, uint64_t cur, uint64_t nm, uint32_t vmask, uint32_t g, uint32_t gpos0,
                                          uint32_t mlo, uint32_t mhi, uint32_t m32, uint32_t a32, uint64_t km1) {
    const uint32_t F0 = (uint32_t)cur, F1 = (uint32_t)(cur >> 32), F2 = (uint32_t)prev, F3 = (uint32_t)(prev >> 32);
    const uint64_t rcc = rc64(cur), rcp = rc64(prev);
    const uint32_t R0 = (uint32_t)rcp, R1 = (uint32_t)(rcp >> 32), R2 = (uint32_t)rcc, R3 = (uint32_t)(rcc >> 32);
    const uint32_t neg_p = 0u - p.P;
    const int C = 128 - 2 * k;
    const uint32_t start31 = gpos0 + 31u - (uint32_t)k + 1u;     // start of the window ending at base 31 of the word (o = 0)
    int o = 0;
#pragma unroll 1
    while (o < 32) {
        // a segment = a run of o over which both extractions read the same three words
        const int wf = o >> 4, ar = C - 2 * o, wr = ar >> 5;
        const int o_end = min((wf + 1) << 4, o + ((ar & 31) >> 1) + 1);
        const uint32_t a0 = wf ? F1 : F0, a1 = wf ? F2 : F1, a2 = wf ? F3 : F2;
        const uint32_t b0 = wr == 0 ? R0 : wr == 1 ? R1 : wr == 2 ? R2 : R3;
        const uint32_t b1 = wr == 0 ? R1 : wr == 1 ? R2 : wr == 2 ? R3 : 0u;
        const uint32_t b2 = wr == 0 ? R2 : wr == 1 ? R3 : 0u;
#pragma unroll 4
        for (; o < o_end; o++) {
            const uint32_t sf = 2u * (uint32_t)o, sr = (uint32_t)(C - 2 * o);   // the funnel shift takes the low 5 bits itself
            const int j = 31 - o;
            uint32_t h_lo = __funnelshift_r(a0, a1, sf), h_hi = 0, r_lo = __funnelshift_r(b0, b1, sr), r_hi = 0;
            if (KBIG) { h_hi = __funnelshift_r(a1, a2, sf) & mhi; r_hi = __funnelshift_r(b1, b2, sr) & mhi; }
            else { h_lo &= mlo; r_lo &= mlo; }
            uint32_t valid = 1;
            if (!CLEAN) {
                valid = (vmask >> j) & 1u;
                uint32_t wn = (uint32_t)((nm >> (31 - j)) & km1);
                if (valid && wn) {
                    if (p.first_pass != 2)
                        pollute_invalid_window(p, ki, k, g, (((uint64_t)h_hi) << 32) | h_lo, (((uint64_t)r_hi) << 32) | r_lo, wn);
                    valid = 0;
                }
            }
            const uint32_t bin = canon_bin<FASTMOD, KBIG>(p, h_lo, h_hi, r_lo, r_hi, m32, a32, neg_p);
            // filter bit = (word umulhi(bin, scale), bit bin & 31); read unconditionally: no branch before the LDS
            uint32_t hit = lds32(sw.f_addr + __umulhi(bin, p.fscale) * 4u) & __funnelshift_l(0u, 1u, bin);   // word & (1 << (bin & 31))
            if (!CLEAN) hit = valid ? hit : 0u;
            // straight-line append: ballot, predicated store, uniform counters; the only branch is the
            // (warp-uniform, rarely taken) drain.  The item holds (bin, o, genome, start of the window at o = 0):
            // the start itself is only needed for the ~22 % of the windows that are drained
            uint32_t m;
            asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; vote.sync.ballot.b32 %0, q, 0xffffffff; }" : "=r"(m) : "r"(hit));
            sts128_if(hit, sw.q_addr + (((sw.qhead + sw.qn + __popc(m & sw.lt_mask)) & (QCAP - 1)) << 4), bin, (uint32_t)o, g, start31);
            sw.qn += __popc(m);
            if (sw.qn >= 32) {
                __syncwarp();
                if (sw.pend) drain_consume(p, sw.pd);
                drain_issue(p, ki, sw.q_addr, sw.qhead, sw.lane, sw.pd);
                sw.pend = true;
                sw.qhead = (sw.qhead + 32) & (QCAP - 1);
                sw.qn -= 32;
                __syncwarp();
            }
        }
    }
}

template <bool FASTMOD, bool KBIG>
__device__ __forceinline__ void scan_pass(const K &p, int ki, int k, ScanWarp &sw, uint32_t first, uint32_t w1, uint32_t stride,
                                          uint32_t warp) {
    const uint64_t km2 = kmask2(k);
    const uint64_t km1 = kmask1(k);
    const uint32_t mlo = (uint32_t)km2, mhi = (uint32_t)(km2 >> 32);
    const uint32_t m32 = (uint32_t)(0x100000000ull / p.P);            // floor(2^32 / P)
    const uint32_t a32 = (uint32_t)(0x100000000ull % p.P);            // 2^32 mod P
    for (uint32_t wb = first + warp * 32; wb < w1; wb += stride) {
        const uint32_t w = wb + sw.lane;
        uint32_t vmask = 0, g = 0, gpos0 = 0;
        uint64_t cur = 0, prev = 0, nm = 0;
        if (w < w1) {
            uint32_t c = p.wordContig[w];
            uint32_t i = w - p.cWordStart[c];
            uint32_t qbase = 32u * i;
            uint32_t nj = min(32u, p.cLen[c] - qbase);
            uint32_t jlo = qbase + 1 >= (uint32_t)k ? 0u : (uint32_t)k - 1 - qbase;   // first end position with a full window
            if (jlo < nj) vmask = (nj == 32 ? 0xFFFFFFFFu : ((1u << nj) - 1u)) & ~((1u << jlo) - 1u);
            g = p.cGenome[c];
            cur = p.seq2[w];
            prev = i ? p.seq2[w - 1] : 0ull;
            nm = ((uint64_t)(i ? p.nmask[w - 1] : 0u) << 32) | p.nmask[w];
            gpos0 = 32u * (w - p.gWordStart[g]);
        }
        // every lane's word entirely inside a contig and free of non-ACGT bytes: no validity work at all
        const bool clean = __all_sync(0xFFFFFFFFu, vmask == 0xFFFFFFFFu && nm == 0);
        if (clean) scan_word<FASTMOD, KBIG, true>(p, ki, k, sw, prev, cur, nm, vmask, g, gpos0, mlo, mhi, m32, a32, km1);
        else scan_word<FASTMOD, KBIG, false>(p, ki, k, sw, prev, cur, nm, vmask, g, gpos0, mlo, mhi, m32, a32, km1);
    }
}

// (one length per CTA -- a single filter load -- was tried and is much slower: the CTAs of the nk lengths then
// work on the same stretch of a genome at the same time and their REDs collide on the same row cache lines)
__global__ void __launch_bounds__(1024, 1) k_scan_filtered(const __grid_constant__ K p, uint32_t w0_, uint32_t w1) {
    extern __shared__ uint32_t s_filter[];
    ScanWarp sw;
    sw.lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    sw.f_addr = (uint32_t)__cvta_generic_to_shared(s_filter);
    sw.q_addr = sw.f_addr + FILTER_BYTES + warp * QCAP * 16u;
    sw.lt_mask = (1u << sw.lane) - 1u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t w0 = w0_ + blockIdx.x * blockDim.x;
    for (int ki = p.kiSplit; ki < p.nk; ki++) {        // (the lengths before kiSplit belong to k_scan_filtered2)
        const int k = p.kmin + ki;
        __syncthreads();
        {   // LDGSTS: every thread has all its 16 B pieces in flight at once, nothing goes through registers
            const uint8_t *src = (const uint8_t *)(p.filter + (size_t)ki * FILTER_WORDS);
            for (uint32_t t = threadIdx.x; t < FILTER_WORDS / 4; t += blockDim.x)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sw.f_addr + t * 16u), "l"(src + (size_t)t * 16u) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        sw.qhead = 0; sw.qn = 0; sw.pend = false;
        sw.pd.e = make_uint2(0u, 0u); sw.pd.meta = 0; sw.pd.g = 0;
        const bool fastmod = k <= 20 && p.P < (1u << 24);
        if (k > 16) {
            if (fastmod) scan_pass<true, true>(p, ki, k, sw, w0, w1, stride, warp);
            else scan_pass<false, true>(p, ki, k, sw, w0, w1, stride, warp);
        } else {
            if (p.P < (1u << 24)) scan_pass<true, false>(p, ki, k, sw, w0, w1, stride, warp);
            else scan_pass<false, false>(p, ki, k, sw, w0, w1, stride, warp);
        }
        __syncwarp();
        if (sw.pend) drain_consume(p, sw.pd);
        if (sw.lane < sw.qn) {   // the remainder of the k pass
            drain_issue(p, ki, sw.q_addr, sw.qhead, sw.lane, sw.pd);
            drain_consume(p, sw.pd);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// k_scan_filtered2: the hot kernel, second formulation.  Same contract as k_scan_filtered (a per-k filter in
// shared memory, filter-positive windows queued per warp, probed 32 at a time, fire-and-forget REDs) with the
// instruction count per window cut down (ncu, config 2: 66 -> see profiles/):
//   * a thread's 32 windows are two runs of 16 over FIXED registers with COMPILE-TIME shift amounts: the reverse-
//     complement stream is shifted once per word (by 66 - 2k bits) so that its windows change source registers at
//     the same place as the forward ones (window o: forward bits [2o, 2o + 2k), reverse bits [62 - 2o, ...)); no
//     segment bookkeeping, no loop counter, no shift registers; two windows are in flight per thread (the
//     "is a batch ready" test runs every second window);
//   * the queue is a plain stack: items (bin, start, genome) are appended at the top and a batch is the TOP 32
//     -- no ring arithmetic, nothing to move (the REDs commute, so the order of the probes is free);
//   * the pass over genome 1 itself (another key table) is a separate instantiation: no run-time branch in the
//     drain.
struct Scan2 {
    uint32_t q_addr;         // shared-window address of this warp's queue (uint4 items)
    uint32_t qn;             // items queued (warp-uniform)
    bool pend;               // a batch of 32 probes is in flight
    Pending pd;
    uint32_t lane, lt_mask;
};

template <bool FIRST>
__device__ __forceinline__ void drain2_consume(const K &p, const Pending &pd) {
    const uint32_t b = pd.meta >> 26;          // bin & 31: all the two tables need of the bin besides the probe word
    const unsigned long long add = ROW_ONE | (pd.meta & (uint32_t)ROW_POS_MASK);
    if (FIRST) {
        const uint32_t slot = pd.e.y + __popc(pd.e.x & ((1u << b) - 1u));
        if (((pd.e.x >> b) & 1u) && slot < p.capSlots) atomicAdd(&p.rows0[slot], add);
    } else {
        const uint32_t id = id_of(p, pd.e, b & 15u);
        if (id != ID_NONE) atomicAdd(&p.rows[(size_t)(pd.g - 1) * p.capIds + id], add);
    }
}

template <bool FIRST>
__device__ __forceinline__ void drain2(const K &p, int ki, Scan2 &sw) {
    __syncwarp();                                            // every lane's appends are visible
    if (sw.pend) drain2_consume<FIRST>(p, sw.pd);
    const uint32_t taken = min(sw.qn, 32u);                  // a full batch but for the flush at the end of a pass
    sw.qn -= taken;
    uint4 it = make_uint4(0u, 0u, 0u, 0u);
    sw.pd.e = make_uint2(0u, 0u);
    if (sw.lane < taken) {
        it = lds128(sw.q_addr + ((sw.qn + sw.lane) << 4));   // x = bin, y = start, z = genome
        sw.pd.e = FIRST ? probe(p, ki, it.x) : probe2(p, ki, it.x);
    }
    sw.pd.meta = (it.y & (uint32_t)ROW_POS_MASK) | ((it.x & 31u) << 26);
    sw.pd.g = it.z;
    sw.pend = true;
    __syncwarp();
}

template <bool FASTMOD, bool KBIG>
__device__ __forceinline__ void window2(const K &p, Scan2 &sw, uint32_t f_addr, uint32_t h_lo, uint32_t h_hi, uint32_t r_lo, uint32_t r_hi,
                                        uint32_t m32, uint32_t a32, uint32_t neg_p, uint32_t start, uint32_t g, bool valid) {
    const uint32_t bin = canon_bin<FASTMOD, KBIG>(p, h_lo, h_hi, r_lo, r_hi, m32, a32, neg_p);
    // filter bit = (word umulhi(bin, scale), bit bin & 31); read unconditionally: no branch before the LDS
    uint32_t hit = lds32(f_addr + __umulhi(bin, p.fscale) * 4u) & __funnelshift_l(0u, 1u, bin);   // word & (1 << (bin & 31))
    if (!valid) hit = 0u;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit != 0u);
    sts128_if(hit, sw.q_addr + ((sw.qn + __popc(m & sw.lt_mask)) << 4), bin, start, g, 0u);
    sw.qn += __popc(m);
}

template <bool FASTMOD, bool KBIG, bool FIRST>
__device__ __forceinline__ void scan2_pass(const K &p, int ki, int k, Scan2 &sw, uint32_t f_addr, uint32_t first, uint32_t w1,
                                           uint32_t stride, uint32_t warp) {
    const uint64_t km2 = kmask2(k);
    const uint64_t km1 = kmask1(k);
    const uint32_t mlo = (uint32_t)km2, mhi = (uint32_t)(km2 >> 32);
    const uint32_t m32 = (uint32_t)(0x100000000ull / p.P);            // floor(2^32 / P)
    const uint32_t a32 = (uint32_t)(0x100000000ull % p.P);            // 2^32 mod P
    const uint32_t neg_p = 0u - p.P;
    const int rs = 66 - 2 * k;                                        // the reverse stream's one shift per word (2 .. 64)
    const int rws = rs >> 5;
    const uint32_t rbs = (uint32_t)rs & 31u;
    for (uint32_t wb = first + warp * 32; wb < w1; wb += stride) {
        const uint32_t w = wb + sw.lane;
        uint32_t vmask = 0, g = 0, gpos0 = 0;
        uint64_t cur = 0, prev = 0, nm = 0;
        if (w < w1) {
            uint32_t c = p.wordContig[w];
            uint32_t i = w - p.cWordStart[c];
            uint32_t qbase = 32u * i;
            uint32_t nj = min(32u, p.cLen[c] - qbase);
            uint32_t jlo = qbase + 1 >= (uint32_t)k ? 0u : (uint32_t)k - 1 - qbase;   // first end position with a full window
            if (jlo < nj) vmask = (nj == 32 ? 0xFFFFFFFFu : ((1u << nj) - 1u)) & ~((1u << jlo) - 1u);
            g = p.cGenome[c];
            cur = p.seq2[w];
            prev = i ? p.seq2[w - 1] : 0ull;
            nm = ((uint64_t)(i ? p.nmask[w - 1] : 0u) << 32) | p.nmask[w];
            gpos0 = 32u * (w - p.gWordStart[g]);
        }
        const uint32_t start31 = gpos0 + 31u - (uint32_t)k + 1u;     // start of the window ending at base 31 of the word (o = 0)
        const uint32_t F0 = (uint32_t)cur, F1 = (uint32_t)(cur >> 32), F2 = (uint32_t)prev, F3 = (uint32_t)(prev >> 32);
        const uint64_t rcc = rc64(cur), rcp = rc64(prev);
        // the stream rc(cur) : rc(prev) shifted right by 66 - 2k bits: the window of o now starts at bit 62 - 2o
        const uint32_t S0 = (uint32_t)rcp, S1 = (uint32_t)(rcp >> 32), S2 = (uint32_t)rcc, S3 = (uint32_t)(rcc >> 32);
        const uint32_t T0 = rws == 0 ? S0 : rws == 1 ? S1 : S2, T1 = rws == 0 ? S1 : rws == 1 ? S2 : S3;
        const uint32_t T2 = rws == 0 ? S2 : rws == 1 ? S3 : 0u, T3 = rws == 0 ? S3 : 0u;
        const uint32_t R0 = __funnelshift_r(T0, T1, rbs), R1 = __funnelshift_r(T1, T2, rbs), R2 = __funnelshift_r(T2, T3, rbs),
                       R3 = __funnelshift_r(T3, 0u, rbs);
        // every lane's word entirely inside a contig and free of non-ACGT bytes: no validity work at all
        const bool clean = __all_sync(0xFFFFFFFFu, vmask == 0xFFFFFFFFu && nm == 0);
#pragma unroll 1
        for (int seg = 0; seg < 2; seg++) {
            const uint32_t a0 = seg ? F1 : F0, a1 = seg ? F2 : F1, a2 = seg ? F3 : F2;
            const uint32_t b0 = seg ? R0 : R1, b1 = seg ? R1 : R2, b2 = seg ? R2 : R3;
            const uint32_t s0 = start31 - 16u * (uint32_t)seg;
            if (clean) {
#pragma unroll
                for (int oo = 0; oo < 16; oo++) {
                    uint32_t h_lo = __funnelshift_r(a0, a1, 2 * oo), h_hi = 0, r_lo = __funnelshift_r(b0, b1, 30 - 2 * oo), r_hi = 0;
                    if (KBIG) { h_hi = __funnelshift_r(a1, a2, 2 * oo) & mhi; r_hi = __funnelshift_r(b1, b2, 30 - 2 * oo) & mhi; }
                    else { h_lo &= mlo; r_lo &= mlo; }
                    window2<FASTMOD, KBIG>(p, sw, f_addr, h_lo, h_hi, r_lo, r_hi, m32, a32, neg_p, s0 - (uint32_t)oo, g, true);
                    if (oo & 1) { while (sw.qn >= 32u) drain2<FIRST>(p, ki, sw); }      // (at most 31 + 2 * 32 items between two tests: the queue holds 96)
                }
            } else {
#pragma unroll 1
                for (int oo = 0; oo < 16; oo++) {
                    const int o = seg * 16 + oo, j = 31 - o;
                    uint32_t h_lo = __funnelshift_r(a0, a1, 2 * oo), h_hi = 0, r_lo = __funnelshift_r(b0, b1, 30 - 2 * oo), r_hi = 0;
                    if (KBIG) { h_hi = __funnelshift_r(a1, a2, 2 * oo) & mhi; r_hi = __funnelshift_r(b1, b2, 30 - 2 * oo) & mhi; }
                    else { h_lo &= mlo; r_lo &= mlo; }
                    bool valid = (vmask >> j) & 1u;
                    const uint32_t wn = (uint32_t)((nm >> (31 - j)) & km1);
                    if (valid && wn) {
                        if (p.first_pass != 2)
                            pollute_invalid_window(p, ki, k, g, (((uint64_t)h_hi) << 32) | h_lo, (((uint64_t)r_hi) << 32) | r_lo, wn);
                        valid = false;
                    }
                    window2<FASTMOD, KBIG>(p, sw, f_addr, h_lo, h_hi, r_lo, r_hi, m32, a32, neg_p, s0 - (uint32_t)oo, g, valid);
                    while (sw.qn >= 32u) drain2<FIRST>(p, ki, sw);
                }
            }
        }
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(1024, 1) k_scan_filtered2(const __grid_constant__ K p, uint32_t w0_, uint32_t w1) {
    extern __shared__ __align__(128) uint32_t s_filter[];
    Scan2 sw;
    sw.lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t f_addr = (uint32_t)__cvta_generic_to_shared(s_filter);
    sw.q_addr = f_addr + FILTER_BYTES + warp * (Q2_ITEMS * 16u);
    sw.lt_mask = (1u << sw.lane) - 1u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t w0 = w0_ + blockIdx.x * blockDim.x;
    // the filter of a length arrives as ONE bulk copy (cp.async.bulk -> UBLKCP) that completes on an mbarrier: the last 16
    // bytes of the shared window -- item 95 of the last warp's queue, which no pass ever reaches (at most 31 + 2 * 32 items)
    const uint32_t bar = f_addr + SCAN2_SMEM - 16u;
    uint32_t parity = 0;
    if (threadIdx.x == 0) bulk::mbar_init(bar, 1u);
    for (int ki = 0; ki < p.kiSplit; ki++) {
        const int k = p.kmin + ki;
        __syncthreads();                                     // every warp is done with the previous filter (and the barrier exists)
        if (threadIdx.x == 0) bulk::load(f_addr, p.filter + (size_t)ki * FILTER_WORDS, FILTER_BYTES, bar);
        bulk::wait(bar, parity);
        parity ^= 1u;
        sw.qn = 0; sw.pend = false;
        sw.pd.e = make_uint2(0u, 0u); sw.pd.meta = 0; sw.pd.g = 0;
        const bool fastmod = k <= 20 && p.P < (1u << 24);
        if (k > 16) {
            if (fastmod) scan2_pass<true, true, FIRST>(p, ki, k, sw, f_addr, w0, w1, stride, warp);
            else scan2_pass<false, true, FIRST>(p, ki, k, sw, f_addr, w0, w1, stride, warp);
        } else {
            if (p.P < (1u << 24)) scan2_pass<true, false, FIRST>(p, ki, k, sw, f_addr, w0, w1, stride, warp);
            else scan2_pass<false, false, FIRST>(p, ki, k, sw, f_addr, w0, w1, stride, warp);
        }
        // the remainder of the k pass
        while (sw.qn) drain2<FIRST>(p, ki, sw);
        __syncwarp();
        if (sw.pend) drain2_consume<FIRST>(p, sw.pd);
        sw.pend = false;
        __syncwarp();
    }
}

// filter bit of every key bin (genome 1's pass: all keys)
__global__ void __launch_bounds__(256) k_build_filter(const __grid_constant__ K p) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)p.nk * p.PW) return;
    uint32_t word = p.br[t].x;
    if (!word) return;
    uint32_t ki = t / p.PW, wi = t - ki * p.PW;
    uint32_t *f = p.filter + (size_t)ki * FILTER_WORDS;
    while (word) {
        int b = __ffs(word) - 1;
        word &= word - 1;
        uint32_t bin = wi * 32u + b;
        atomicOr(&f[__umulhi(bin, p.fscale)], 1u << (bin & 31u));
    }
}

// k_build_at: the key table of genomes 2..G (see id_of) from the keys that survived genome 1, and -- keys that
// are dead already need no filter bit: fewer false drains -- the filter of the alive keys.  One thread per
// 16 bins.
__device__ __forceinline__ uint64_t window_at(const K &p, uint32_t g, uint32_t lp, int k);   // defined with k_alive1 below

// slot -> id of one key that survived genome 1 (its rank in dict order: longer k-mers ending at the same place come
// first), and what the id-ordered kernels need per id
__device__ __forceinline__ uint32_t assign_id(const K &p, uint32_t slot, int ki) {
    const int k = p.kmin + ki;
    const uint32_t start = (uint32_t)(p.rows0[slot] & ROW_POS_MASK);
    const uint32_t end = start + (uint32_t)k - 1u;
    const uint32_t m = p.emitMask[end];
    const uint32_t id = p.emitOff[end] + __popc(ki < 31 ? (m >> (ki + 1)) : 0u);
    if (id >= p.capIds) {            // more ids than the buffers hold: the host grows them and repeats the run
        atomicOr(&p.cnt[C_FLAGS], OVF_IDS);
        p.remap[slot] = ID_NONE;
        return ID_NONE;
    }
    const uint64_t key = window_at(p, 0, start, k);
    p.remap[slot] = id;
    p.key1[id] = key;
    p.idMeta[id] = start | ((uint32_t)ki << 26);
    p.alive[id] = 1;
    // genome 1: every key sits on its (+) strand; a palindrome matches both scans and the later '-' match
    // overwrites (getCandidateKmers.py:181-187)
    p.pos32[id] = start | (((rc64(key) >> (64 - 2 * k)) == key) ? 0x80000000u : 0u);
    return id;
}

// (with more than one genome the ids are assigned here as well -- every slot lies in exactly one 16-bin word --
// instead of in a separate k_assign_ids pass)
__global__ void __launch_bounds__(256) k_build_at(const __grid_constant__ K p, int with_filter) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)p.nk * p.PW16) return;
    const uint32_t ki = t / p.PW16, wi = t - ki * p.PW16;
    uint2 out = make_uint2(0u, 0u);
    if ((wi >> 1) < p.PW) {
        const uint2 e = p.br[(size_t)ki * p.PW + (wi >> 1)];
        const uint32_t half = (wi & 1u) ? (e.x >> 16) : (e.x & 0xFFFFu);
        if (half) {
            uint32_t slot = e.y + ((wi & 1u) ? __popc(e.x & 0xFFFFu) : 0u);
            uint32_t bm = 0, n = 0, id0 = 0, id1 = 0;
            bool wide = false;
            for (uint32_t m = half; m; m &= m - 1, slot++) {
                if (slot >= p.capSlots || p.remap[slot] == ID_NONE) continue;         // k_alive1: not unique in genome 1
                const uint32_t id = assign_id(p, slot, (int)ki);
                if (id == ID_NONE) continue;
                bm |= 1u << (__ffs(m) - 1);
                if (n == 0) id0 = id; else if (n == 1) id1 = id;
                wide |= id >= p.atWide;
                n++;
            }
            if (n > 2 || wide) {
                const uint32_t idx = atomicAdd(p.ovfCount, 1u);
                if (idx < p.ovfCap) {
                    slot = e.y + ((wi & 1u) ? __popc(e.x & 0xFFFFu) : 0u);
                    for (uint32_t m = half; m; m &= m - 1, slot++)
                        p.ovf[(size_t)idx * 16u + (__ffs(m) - 1)] = slot < p.capSlots ? p.remap[slot] : ID_NONE;
                } else {
                    atomicOr(&p.cnt[C_FLAGS], OVF_LISTS);
                    bm = 0; n = 0;
                }
                id0 = AT_SLOW; id1 = idx;
            }
            if (n) out = make_uint2(bm | (id0 << 16), (id0 >> 16) | (id1 << 8));
            if (with_filter) {
                uint32_t *f = p.filter + (size_t)ki * FILTER_WORDS;
                for (uint32_t m = bm; m; m &= m - 1) {
                    const uint32_t bin = wi * 16u + (__ffs(m) - 1);
                    atomicOr(&f[__umulhi(bin, p.fscale)], 1u << (bin & 31u));
                }
            }
        }
    }
    p.at[t] = out;
}

// ------------------------------------------------------------------------------------------
// k_junction: windows of the '~'-joined strings that touch a junction (getCandidateKmers.py:99-100):
// they never become k-mers (:40-43) but khmer counts them ('~' hashes like G).  One thread per
// (contig boundary, k, strand string, window offset) in the virtual joined coordinates.
__global__ void __launch_bounds__(128) k_junction(const __grid_constant__ K p, uint32_t c0, uint32_t c1) {
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
    pollute(p, ki, g, h < r ? h : r, rev != 0);
}

// ------------------------------------------------------------------------------------------
// zero fill on the SMs.  cudaMemsetAsync may run on a copy engine, where it queues up behind the 5 MB H2D
// chunks of a pipelined run (measured: up to 0.1 ms per memset); a kernel does not.
struct ZeroArgs {            // up to four buffers per launch
    uint4 *p[4];
    unsigned long long n16[4];
    int cnt;
};

__global__ void __launch_bounds__(256) k_zero(const ZeroArgs a) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int b = 0; b < a.cnt; b++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n16[b]; i += (size_t)gridDim.x * blockDim.x) a.p[b][i] = z;
}

// the first n entries (n = min(*n_dev, cap): a count only the device knows) of each of `nrows` rows of `pitch`
// 8-byte words: the per-genome rows are allocated by capacity and zeroed by need
__global__ void __launch_bounds__(256) k_zero_rows(unsigned long long *base, uint32_t pitch, uint32_t nrows, const uint32_t *n_dev, uint32_t cap) {
    const uint32_t n = min(*n_dev, cap);
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 2u;     // pitch is a multiple of 256: rows stay 16-byte aligned
    if (i >= n) return;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t r = blockIdx.y; r < nrows; r += gridDim.y) *(uint4 *)(base + (size_t)r * pitch + i) = z;
}

// ------------------------------------------------------------------------------------------
// bits = seen once and not seen again
__global__ void __launch_bounds__(256) k_candbits(const __grid_constant__ K p) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)p.nk * p.PW) return;
    uint32_t lo = p.sk[2 * (size_t)t], hi = p.sk[2 * (size_t)t + 1];
    p.br[t].x = ((lo & ~(lo >> 16)) & 0xFFFFu) | ((hi & ~(hi >> 16)) << 16);
}

// device-wide exclusive scan in two launches (tile sums, then every tile sums the tile sums before it).  MODE: the value of element x is in[x * S] (SV_U32), its
// popcount (SV_POPC) or the byte in[x] (SV_U8); the prefix goes to out[x * S] (S = 2 for the interleaved
// bitmap / rank table)
enum { SV_U32 = 0, SV_POPC = 1, SV_U8 = 2 };
template <int MODE, int S>
__device__ __forceinline__ uint32_t scan_val(const void *in, uint32_t x) {
    if (MODE == SV_U8) return ((const uint8_t *)in)[x];
    uint32_t v = ((const uint32_t *)in)[(size_t)x * S];
    return MODE == SV_POPC ? __popc(v) : v;
}

template <int MODE, int S>
__global__ void __launch_bounds__(256) k_scan_p1(const void *in, uint32_t n, uint32_t *blockSums) {
    __shared__ uint32_t sm[8];
    uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
    for (int t = 0; t < SCAN_TILE / 256; t++) {
        uint32_t x = base + t * 256 + threadIdx.x;
        if (x < n) s += scan_val<MODE, S>(in, x);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t tot = 0; for (int i = 0; i < 8; i++) tot += sm[i]; blockSums[blockIdx.x] = tot; }
}

template <int MODE, int S>
__global__ void __launch_bounds__(256) k_scan_p3(const void *in, uint32_t *out, uint32_t n, const uint32_t *blockSums,
                                                 uint32_t *total, uint32_t *total_host) {
    __shared__ uint32_t sm[8];
    __shared__ uint32_t sb[8];
    // the tile's base = sum of the tile sums before it (a few thousand L2-resident values at most: cheaper than a
    // single-CTA pass over them between the two launches)
    uint32_t part = 0;
    for (uint32_t x = threadIdx.x; x < blockIdx.x; x += 256) part += blockSums[x];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = part;
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * (SCAN_TILE / 256);
    uint32_t v[SCAN_TILE / 256];
    uint32_t s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_TILE / 256; t++) {
        uint32_t x = base + t;
        v[t] = x < n ? scan_val<MODE, S>(in, x) : 0;
        s += v[t];
    }
    uint32_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) sm[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0, tile_base = 0;
    for (int i = 0; i < (int)(threadIdx.x >> 5); i++) woff += sm[i];
    for (int i = 0; i < 8; i++) tile_base += sb[i];
    uint32_t run = tile_base + woff + incl - s;
#pragma unroll
    for (int t = 0; t < SCAN_TILE / 256; t++) {
        uint32_t x = base + t;
        if (x < n) out[(size_t)x * S] = run;
        run += v[t];
    }
    // the total also goes to host-mapped memory: the host reads it after a stream synchronise, with no device ->
    // host copy that would queue up behind the H2D copies of a pipelined run
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) { *total = run; if (total_host) *total_host = run; }
}

// ------------------------------------------------------------------------------------------
// k_alive: per slot, the reference's rules for every genome on this context.
//   the bin must hold exactly one window of the genome (khmer get() == 1, :35-38) and that window
//   must be the key itself (read back from the packed genome at the recorded start);
//   occurrence on the forward string ('+'): needs count == 1 in fkh  -> no JF polluter
//   occurrence on the rc string ('-')     : needs count == 1 in rkh  -> no JR polluter
//   palindrome, other genomes: in F ^ R  <=>  exactly one of the two tables reads 1 (:62-63)
//   palindrome, first genome : in F - R  <=>  fkh reads 1 and rkh does not      (:58-59)
__device__ __forceinline__ uint64_t window_at(const K &p, uint32_t g, uint32_t lp, int k) {
    uint32_t w = p.gWordStart[g] + (lp >> 5), o = lp & 31;
    uint64_t a = p.seq2[w];
    uint64_t x = o ? ((a << (2 * o)) | (p.seq2[w + 1] >> (64 - 2 * o))) : a;   // k <= 32 bases from offset o
    return x >> (64 - 2 * k);
}

// returns 0 dead, 1 alive on '+', 2 alive on '-'; h = the window the genome holds at the row's start (read by the
// caller, and only when the row's count is 1: otherwise the start field is a sum of starts)
__device__ __forceinline__ int row_state_h(unsigned long long v, uint64_t h, uint64_t key_fwd, uint64_t key_rc, bool pal, bool first) {
    if (((v >> 32) & ROW_CNT_MASK) != 1) return 0;
    bool jf = (v & ROW_JF) != 0, jr = (v & ROW_JR) != 0;
    if (pal) return (h == key_fwd && (first ? (!jf && jr) : (jf != jr))) ? 2 : 0;   // later '-' match overwrites (:181-187)
    if (h == key_fwd) return jf ? 0 : 1;
    if (h == key_rc) return jr ? 0 : 2;
    return 0;   // the single occupant is a different k-mer: a sketch collision with the key absent
}

__device__ __forceinline__ int row_state(const K &p, uint32_t g, unsigned long long v, uint64_t key_fwd, uint64_t key_rc,
                                         int k, bool pal, bool first) {
    if (((v >> 32) & ROW_CNT_MASK) != 1) return 0;
    uint64_t h = window_at(p, g, (uint32_t)(v & ROW_POS_MASK), k);
    bool jf = (v & ROW_JF) != 0, jr = (v & ROW_JR) != 0;
    if (pal) return (h == key_fwd && (first ? (!jf && jr) : (jf != jr))) ? 2 : 0;   // later '-' match overwrites (:181-187)
    if (h == key_fwd) return jf ? 0 : 1;
    if (h == key_rc) return jr ? 0 : 2;
    return 0;   // the single occupant is a different k-mer: a sketch collision with the key absent
}

// length of the k-mer a slot holds: slots are numbered per k, in bin order
__device__ __forceinline__ int slot_ki(const uint32_t *s_base, int nk, uint32_t idx) {
    int ki = 0;
    while (ki + 1 < nk && s_base[ki + 1] <= idx) ki++;
    return ki;
}

// k_alive1: genome 1's verdict per slot (bin order); every survivor sets its bit at (end position, k), from
// which a popcount scan over genome 1's end positions derives its id = its rank in the reference's dict
// order (genome-1 end position ascending, longest first: getCandidateKmers.py:171-187 + pyahocorasick's
// emission order)
__global__ void __launch_bounds__(256) k_alive1(const __grid_constant__ K p) {
    __shared__ uint32_t s_base[32];
    if (threadIdx.x < (unsigned)p.nk) s_base[threadIdx.x] = p.br[(size_t)threadIdx.x * p.PW].y;
    __syncthreads();
    const uint32_t nslots = p.cnt[C_NSLOTS];
    if (nslots > p.capSlots && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&p.cnt[C_FLAGS], OVF_SLOTS);
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= min(nslots, p.capSlots)) return;
    const int ki = slot_ki(s_base, p.nk, (uint32_t)idx);
    const int k = p.kmin + ki;
    unsigned long long v0 = p.rows0[idx];
    bool ok = ((v0 >> 32) & ROW_CNT_MASK) == 1;
    if (ok) {
        // the key as it reads on genome 1's (+) strand: the single occupant of its bin there
        const uint32_t start = (uint32_t)(v0 & ROW_POS_MASK);
        uint64_t key = window_at(p, 0, start, k);
        uint64_t rc = rc64(key) >> (64 - 2 * k);
        bool pal = key == rc;
        ok = row_state(p, 0, v0, key, rc, k, pal, true) == (pal ? 2 : 1);
        if (ok) atomicOr(&p.emitMask[start + (uint32_t)k - 1u], 1u << ki);
    }
    p.remap[idx] = ok ? 0u : ID_NONE;
}

// k_assign_ids: slot -> id for the survivors of genome 1, and what the id-ordered kernels need per id
__global__ void __launch_bounds__(256) k_assign_ids(const __grid_constant__ K p) {
    __shared__ uint32_t s_base[32];
    if (threadIdx.x < (unsigned)p.nk) s_base[threadIdx.x] = p.br[(size_t)threadIdx.x * p.PW].y;
    __syncthreads();
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= min(p.cnt[C_NSLOTS], p.capSlots) || p.remap[idx] == ID_NONE) return;
    assign_id(p, (uint32_t)idx, slot_ki(s_base, p.nk, (uint32_t)idx));
}

// multi-GPU, genome 1's own pass shared out by position: every rank's rows0 holds the windows of ITS share.  What
// the rules need of a row is "exactly one window in the bin, and where": count saturated at 3 and the start of a
// lone window fit 32 bits whose plain SUM over the ranks keeps both (count in bits 26.., a carry out of the start
// field can only occur when two ranks contributed, i.e. when the count says >= 2 anyway).  Half the bytes of the
// 64-bit rows on the wire; the polluter flags are not exchanged: every rank re-derives them (k_scan_plain<true>,
// k_junction), they are few.
__global__ void __launch_bounds__(256) k_rows0_pack(const __grid_constant__ K p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.capSlots) return;
    const unsigned long long v = p.rows0[i];
    const uint32_t c = (uint32_t)min((unsigned long long)3, (v >> 32) & ROW_CNT_MASK);
    p.rows0p[i] = (c << 26) | (c == 1 ? (uint32_t)(v & ROW_POS_MASK) : 0u);
}

__global__ void __launch_bounds__(256) k_rows0_unpack(const __grid_constant__ K p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.capSlots) return;
    const uint32_t v = p.rows0p[i];
    const uint32_t c = v >> 26;
    p.rows0[i] = ((unsigned long long)min(c, 2u) << 32) | (c == 1 ? (v & (uint32_t)ROW_POS_MASK) : 0u);
}

// ---- exchanges over peer memory.  A producer signals by writing the run's number into ITS slot of every peer's
// flag table (after a system-wide fence; the kernel boundary before it has made the data visible); a consumer
// waits until every slot of its own table has reached that number, then reads the peers' arenas with loads that
// bypass L1 (peer memory is cached at its home, not here).  The wait is bounded (~9 s): a rank that never signals ends in an error, not in a hang.
// both in one launch (a producer's signal is always followed by its own wait for the others)
__global__ void k_xsync(const __grid_constant__ K p, int phase) {
    if (threadIdx.x < p.nRanks) {
        __threadfence_system();
        *((volatile uint32_t *)(p.peer[threadIdx.x]) + phase * 16 + p.rank) = p.xEpoch;
        const volatile uint32_t *f = (const volatile uint32_t *)(p.peer[p.rank]) + phase * 16 + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(*f - p.xEpoch) < 0) {
            if (clock64() - t0 > (1ll << 34)) { atomicOr(&p.cnt[C_FLAGS], OVF_PEER_TIMEOUT); break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
}

__global__ void k_xwait(const __grid_constant__ K p, int phase, uint32_t value) {
    if (threadIdx.x < p.nRanks) {
        const volatile uint32_t *f = (const volatile uint32_t *)(p.peer[p.rank]) + phase * 16 + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(*f - value) < 0) {
            if (clock64() - t0 > (1ll << 34)) { atomicOr(&p.cnt[C_FLAGS], OVF_PEER_TIMEOUT); break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

// packed rows, stage A: this rank sums ITS slice of the slots over all ranks' shares; stage B: every rank
// collects all slices (each from the rank that summed it) and goes back to the row format
__device__ __forceinline__ void slice_of(uint32_t n, uint32_t nRanks, uint32_t r, uint32_t &a, uint32_t &b) {
    const uint32_t per = ((n + nRanks - 1) / nRanks + 3u) & ~3u;
    a = min(n, per * r); b = min(n, a + per);
}

__global__ void __launch_bounds__(256) k_rows0_reduce_slice(const __grid_constant__ K p) {
    uint32_t a, b;
    slice_of(p.capSlots, p.nRanks, p.rank, a, b);          // (slices start and end at multiples of 4 slots; so does capSlots)
    const uint32_t i = a + 4u * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= b) return;
    uint4 s = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t r = 0; r < p.nRanks; r++) {
        const uint4 x = __ldcg((const uint4 *)(p.peer[r] + p.xRows0p + 4ull * i));
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    *(uint4 *)(p.peer[p.rank] + p.xRows0s + 4ull * i) = s;
}

__device__ __forceinline__ unsigned long long unpack_row(uint32_t v) {
    const uint32_t c = v >> 26;
    return ((unsigned long long)min(c, 2u) << 32) | (c == 1 ? (v & (uint32_t)ROW_POS_MASK) : 0u);
}

__global__ void __launch_bounds__(256) k_rows0_gather_unpack(const __grid_constant__ K p) {
    const uint32_t i = 4u * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= p.capSlots) return;
    const uint32_t per = ((p.capSlots + p.nRanks - 1) / p.nRanks + 3u) & ~3u;
    const uint4 v = __ldcg((const uint4 *)(p.peer[i / per] + p.xRows0s + 4ull * i));
    ulonglong2 lo, hi;
    lo.x = unpack_row(v.x); lo.y = unpack_row(v.y); hi.x = unpack_row(v.z); hi.y = unpack_row(v.w);
    *(ulonglong2 *)(p.rows0 + i) = lo;
    *(ulonglong2 *)(p.rows0 + i + 2) = hi;
}

// one byte per id over all ranks: MIN (alive) / MAX (pick), four ids per thread.  In place: a peer that reads this
// rank's bytes while they are being replaced by the merged ones sees either, and both give it the same result
template <bool IS_MAX>
__global__ void __launch_bounds__(256) k_bytes_merge(const __grid_constant__ K p, uint32_t off) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 4u >= p.capIds) return;
    uint32_t v = IS_MAX ? 0u : 0xFFFFFFFFu;
    for (uint32_t r = 0; r < p.nRanks; r++) {
        const uint32_t x = __ldcg((const uint32_t *)(p.peer[r] + off) + i);
        v = IS_MAX ? __vmaxu4(v, x) : __vminu4(v, x);
    }
    ((uint32_t *)(p.peer[p.rank] + off))[i] = v;
}

// k_aliveN: the rules for one (id, genome >= 2) per thread, for the genomes [ga, gb) of this context; ids follow
// genome 1's coordinates, so the rows stream and -- for genomes that are largely collinear with genome 1 -- so do
// the read-backs.  The verdict goes to alive[], the place and strand to pos32 (what the output and k_pick read:
// half the bytes of the rows, which are not touched again).
#ifndef PFB_ALIVE_GPT
#define PFB_ALIVE_GPT 4
#endif
constexpr uint32_t ALIVE_GPT = PFB_ALIVE_GPT;   // genomes per thread: independent row -> genome-word load chains in flight

__global__ void __launch_bounds__(256) k_aliveN(const __grid_constant__ K p, uint32_t ga, uint32_t gb) {
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= min(p.cnt[C_NIDS], p.capIds)) return;
    const uint32_t g0 = ga + blockIdx.y * ALIVE_GPT;
    const int k = p.kmin + (int)(p.idMeta[id] >> 26);
    const uint64_t key = p.key1[id];
    const uint64_t rc = rc64(key) >> (64 - 2 * k);
    unsigned long long v[ALIVE_GPT];
    uint64_t h[ALIVE_GPT];
#pragma unroll
    for (uint32_t i = 0; i < ALIVE_GPT; i++) v[i] = g0 + i < gb ? p.rows[(size_t)(g0 + i - 1) * p.capIds + id] : 0ull;
#pragma unroll
    for (uint32_t i = 0; i < ALIVE_GPT; i++)
        h[i] = ((v[i] >> 32) & ROW_CNT_MASK) == 1 ? window_at(p, g0 + i, (uint32_t)(v[i] & ROW_POS_MASK), k) : 0ull;
#pragma unroll
    for (uint32_t i = 0; i < ALIVE_GPT; i++) {
        if (g0 + i >= gb) break;
        const int st = row_state_h(v[i], h[i], key, rc, key == rc, false);
        if (st == 0) p.alive[id] = 0;
        // (0xFFFFFFFF: not a shared k-mer as far as this genome is concerned)
        p.pos32[(size_t)(g0 + i) * p.capIds + id] = st ? ((uint32_t)(v[i] & ROW_POS_MASK) | (st == 2 ? 0x80000000u : 0u)) : 0xFFFFFFFFu;
    }
}

// k_id_records: Tm, GC count, length and filter flags of every id (functions of the key alone: computed as soon
// as the ids exist, so that they can travel to the host while the other genomes are still being scanned)
__global__ void __launch_bounds__(256) k_id_records(const __grid_constant__ K p) {
    __shared__ uint32_t s_nn[16];
    if (threadIdx.x < 16) s_nn[threadIdx.x] = c_nn[threadIdx.x];
    __syncthreads();
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= min(p.cnt[C_NIDS], p.capIds)) return;
    const int k = p.kmin + (int)(p.idMeta[id] >> 26);
    const uint64_t h = p.key1[id];
    double tm;
    uint32_t f = filter_flags(h, k, p, s_nn, &tm);
    if ((rc64(h) >> (64 - 2 * k)) == h) f |= PFB_F_PALIN;
    p.idTm[id] = tm;
    p.idInfo[id] = (uint16_t)((uint32_t)(k - 1) | ((uint32_t)gc_count(h) << 5) | (f << 11));
}

// k_pick: the per-position pick (getCandidateKmers.py:222-247, 376-388): a k-mer is picked when, in at least one
// genome, no earlier k-mer that passes the filters sits at the same (contig, start).
// Group mates need no sort and no per-position array: two shared k-mers that start at the same place of
// some genome read the same sequence there, so one is a prefix of the other ('+' strand) or a suffix of it
// ('-' strand), and as both are unique in genome 1 they share their START (resp. their END) in genome 1.
// The only possible earlier mates of record (start1, k) are therefore the records (same end1, k' > k) and
// (same start1, k' < k), which the emit bitmap locates directly; whether such a record really is a mate
// in genome g is decided by comparing the two places in g.  One thread per id, loop over genomes: every
// access of the loop is coalesced (places are in id order).
struct Mates {
    static constexpr int CAP = 8;
    uint32_t id[CAP];
    int n;                  // > CAP: overflow, re-enumerate per genome
};

__device__ __forceinline__ bool id_passes(const K &p, uint32_t id) {
    return p.alive[id] && ((p.idInfo[id] >> 11) & 7u) == 7u;
}

// visits the earlier potential mates of (start1, ki); f(mate id) returns true to stop
template <typename F>
__device__ __forceinline__ bool for_each_mate(const K &p, uint32_t start1, int ki, F f) {
    const uint32_t end1 = start1 + (uint32_t)(p.kmin + ki) - 1u;
    const uint32_t m = p.emitMask[end1], off = p.emitOff[end1];
    for (uint32_t hi = ki < 31 ? (m >> (ki + 1)) : 0u, kj = ki + 1; hi; hi >>= 1, kj++)        // same end, longer
        if ((hi & 1u) && f(off + __popc(kj < 31 ? (m >> (kj + 1)) : 0u))) return true;
    for (int kj = 0; kj < ki; kj++) {                                                            // same start, shorter
        const uint32_t e2 = start1 + (uint32_t)(p.kmin + kj) - 1u;
        const uint32_t m2 = p.emitMask[e2];
        if (((m2 >> kj) & 1u) && f(p.emitOff[e2] + __popc(m2 >> (kj + 1)))) return true;
    }
    return false;
}

__global__ void __launch_bounds__(256) k_pick(const __grid_constant__ K p) {
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= min(p.cnt[C_NIDS], p.capIds) || !id_passes(p, (uint32_t)id)) return;
    const uint32_t meta = p.idMeta[id];
    const uint32_t start1 = meta & (uint32_t)ROW_POS_MASK;
    const int ki = (int)(meta >> 26);
    Mates mt;
    mt.n = 0;
    for_each_mate(p, start1, ki, [&](uint32_t mid) {
        if (mid < p.capIds && id_passes(p, mid)) {
            if (mt.n < Mates::CAP) mt.id[mt.n] = mid;
            mt.n++;
        }
        return false;
    });
    bool picked = mt.n == 0;
    const bool overflow = mt.n > Mates::CAP;
    // the places of four genomes are requested before the first is used: four times the bytes in flight per thread
    for (uint32_t g0 = 0; g0 < p.nGenomes && !picked; g0 += 4) {
        uint32_t lp[4];
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) lp[i] = g0 + i < p.nGenomes ? p.pos32[(size_t)(g0 + i) * p.capIds + id] & 0x7FFFFFFFu : 0u;
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) {
            if (g0 + i >= p.nGenomes || picked) continue;
            const uint32_t *row = p.pos32 + (size_t)(g0 + i) * p.capIds;
            bool first = true;
            if (overflow) {
                first = !for_each_mate(p, start1, ki, [&](uint32_t mid) {
                    return mid < p.capIds && id_passes(p, mid) && (row[mid] & 0x7FFFFFFFu) == lp[i];
                });
            } else {
                for (int j = 0; j < mt.n; j++)
                    if ((row[mt.id[j]] & 0x7FFFFFFFu) == lp[i]) first = false;
            }
            picked = first;
        }
    }
    if (picked) p.pick[id] = 1;
}

// the end of a run in ONE launch: (peer mode) pick = MAX over the ranks' arenas; the counts of shared and picked ids; the
// block that finishes last publishes the counters to host-mapped memory and (peer mode) tells the peers that this rank
// has read everything it needed from their arenas
template <bool PEER>
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ K p, uint32_t *host_counts) {
    __shared__ bool last;
    const uint32_t n = min(p.cnt[C_NIDS], p.capIds);
    const uint32_t n4 = (p.capIds + 3u) / 4u;
    uint32_t a = 0, b = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        uint32_t pk;
        if (PEER) {
            pk = 0u;
            for (uint32_t r = 0; r < p.nRanks; r++) pk = __vmaxu4(pk, __ldcg((const uint32_t *)(p.peer[r] + p.xPick) + i));
            ((uint32_t *)p.pick)[i] = pk;
        } else {
            pk = ((const uint32_t *)p.pick)[i];
        }
        const uint32_t left = n > 4u * i ? n - 4u * i : 0u;                        // ids of this word below the id count
        const uint32_t m = (left >= 4u ? 0xFFFFFFFFu : (1u << (8u * left)) - 1u) & 0x01010101u;
        a += __popc(((const uint32_t *)p.alive)[i] & m);
        b += __popc(pk & m);
    }
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&p.cnt[C_NALIVE], a);
        if (b) atomicAdd(&p.cnt[C_NPICK], b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&p.cnt[C_TICKET], 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        if (threadIdx.x < 8) host_counts[threadIdx.x] = ((volatile uint32_t *)p.cnt)[threadIdx.x];
        if (PEER && threadIdx.x < p.nRanks) {
            __threadfence_system();
            *((volatile uint32_t *)(p.peer[threadIdx.x]) + XP_DONE * 16 + p.rank) = p.xEpoch;
        }
    }
}
__global__ void k_oligo(const __grid_constant__ K p, uint64_t h, int k, double *out) {
    __shared__ uint32_t s_nn[16];
    for (int i = 0; i < 16; i++) s_nn[i] = c_nn[i];
    int gc = gc_count(h);
    out[0] = oligo_tm(h, k, gc, p, s_nn);
    out[1] = __dmul_rn(__ddiv_rn((double)gc, (double)k), 100.0);
}

// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct HostGenome {
    const uint8_t *bases;        // contigs concatenated (pfb_add_genome) or the FASTA file image (pfb_add_fasta)
    uint64_t n_bases;            // bases; for a FASTA genome known once the device has parsed it
    std::vector<uint64_t> off;   // n_contigs + 1
    int fasta = 0;               // 0: bases given, fa::FMT_FASTA / fa::FMT_GENBANK: file image, parsed on the device
    uint64_t file_bytes = 0;
    std::vector<uint64_t> hdr_off;   // file image: offset of every record's '>' / LOCUS line in the file
};

// one batch of FASTA files being parsed on the device
struct FaBatch {
    size_t g0 = 0, g1 = 0;       // genomes [g0, g1)
    uint32_t nTiles = 0, cap = 0;
    fa::Args args{};
    cudaEvent_t done = nullptr;
    bool pending = false;
};

}  // namespace

// ------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen): a single-GPU user needs no NCCL at all, a torch process gets the copy
// torch has already loaded (same soname), anything else the system library.
namespace nccl {
struct Comm;
struct UniqueId { char internal[128]; };
enum { SUM = 0, MAX = 2, MIN = 3 };
enum { U8 = 1, U32 = 3 };
struct Api {
    void *handle = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm **, int, UniqueId, int) = nullptr;
    int (*CommInitAll)(Comm **, int, const int *) = nullptr;
    int (*CommDestroy)(Comm *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
}  // namespace nccl

struct pfb_multi;

struct pfb_ctx {
    int device = 0;
    int num_sms = 148;
    int scan_mode = 0;   // 0 auto, 1 plain L2 probe, 2 shared-memory filter
    int scan_kernel = 2; // which formulation of the filtered scan: 2 (default) or 1 (PFB200_SCAN_KERNEL=1, the first one; A/B)
    pfb_params prm{};
    std::string err;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;          // H2D copies of pfb_run run here, overlapped with compute
    cudaStream_t d2hStream = nullptr;           // result copies of pfb_run run here, overlapped with compute
    cudaStream_t sideStream = nullptr;          // packing of genomes 2..G runs here, beside genome 1's pipeline of small kernels
    std::vector<cudaEvent_t> evEnc;             // one per batch: packed
    cudaEvent_t evSide = nullptr;
    std::vector<cudaEvent_t> evCopy;            // one per genome: its bases have arrived
    std::vector<cudaEvent_t> evOut;             // result pieces ready for their device -> host copy
    bool pipelined = false;
    // PFB200_TRACE=1: timestamps (ms after the first H2D copy was queued) of the pipeline's milestones on stderr
    volatile uint32_t *countsHost = nullptr;    // mapped pinned: the C_* counters as published by the last run
    uint32_t *countsDev = nullptr;              // the same memory as the device sees it
    uint32_t at_wide = 0xFFFFFFu;               // PFB200_AT_WIDE lowers it so that tests reach the id-list path
    bool fterm_uploaded = false;
    bool tracing = false;
    std::vector<std::pair<std::string, cudaEvent_t>> trace;
    // FASTA ingest: file images are staged in `fa`, parsed on parseStream, record tables land in mapped host memory
    cudaStream_t parseStream = nullptr;
    bool any_fasta = false;
    size_t laid_out = 0;                        // genomes covered by the current layout (FASTA: 1, then all)
    std::vector<uint64_t> faStart;
    FaBatch faBatch[2];
    void *faHost[2] = {nullptr, nullptr};       // cudaHostAlloc(Mapped)
    size_t faHostCap[2] = {0, 0};
    std::vector<uint64_t> fa64host[2];
    std::vector<uint32_t> fa32host[2];
    std::vector<uint32_t> t32host, t32hostb;    // contig / genome tables (kept alive for the async copy)
    std::vector<uint64_t> t64host, t64hostb;
    uint64_t totalFa = 0, fa_launches = 0;
    cudaEvent_t ev[12]{};
    cudaEvent_t evX[4][2]{};                    // around each collective of a run
    bool xUsed[4] = {false, false, false, false};
    std::vector<HostGenome> genomes;
    // layout (host)
    std::vector<uint32_t> cWordStart, cLen, cGenome, cVirt, gWordStart, gFirstContig, gNumContigs, gVirtLen;
    std::vector<uint64_t> cRawStart, gRawStart;
    uint64_t totalRaw = 0;
    uint32_t totalWords = 0;
    bool uploaded = false, computed = false;
    // capacities (see K): `cap*` is what the kernels of the next run are given, `alloc*` what the buffers hold;
    // `last*` the counts of the last completed run (0: none yet)
    uint32_t capSlots = 0, capIds = 0, capCand = 0, capCandMin = 0;
    uint64_t lastSlots = 0, lastIds = 0;
    double cap_share = 0.0;                     // PFB200_CAP_SHARE (tests): a-priori share of genome 1's windows the buffers are sized for
    int reruns = 0;                             // runs repeated because a capacity was too small (all runs of the context)
    int caps_for_prefilter = -1;
    uint64_t caps_for_windows = 0;
    // multi-GPU
    nccl::Comm *comm = nullptr;
    bool comm_owned = false;
    int nRanks = 1, rank = 0;
    int shard_g1 = 1;                           // share genome 1's own pipeline out over the ranks (PFB200_SHARD_G1=0 disables)
    bool sharded_now = false;                   // ... and whether the current run does
    int fetch_records = 1;                      // 0: the per-id records stay on the device (ranks > 0 of a job: identical everywhere)
    // exchanges over peer memory (see K::peer): 0 = NCCL collectives, 1 = peer arenas
    int xchg_pref = 1;                          // PFB200_XCHG=nccl -> 0
    bool peer_now = false;                      // the current run exchanges through the arenas
    bool x_dirty = false;                       // the arena was (re)allocated: the ranks must map each other's again
    DevBuf xa, xh;                              // the arena; staging of the IPC handles
    uint8_t *xPeer[16] = {};                    // mapped arenas of the other ranks (IPC) / their pointers (one process)
    bool xOpened[16] = {};
    uint32_t xCaps[3] = {0, 0, 0};              // capCand, capSlots, capIds the arena was laid out for
    uint32_t epoch = 0;                         // runs (attempts) on this arena
    int emulate_ranks = 0;                      // > 1 (tests): ONE context plays that many ranks of the shared-out genome-1
                                                // pipeline one after the other (lists, table update, packed rows: all but NCCL)
    int eff_ranks() const { return emulate_ranks > 1 ? emulate_ranks : nRanks; }
    pfb_multi *owner = nullptr;
    // results on the host (pinned, indexed by id, rows `hostPitch` apart)
    bool stream_out = false;                    // the run copies results out as they become final
    bool host_valid = false;
    uint64_t hostPitch = 0;
    DevBuf hWord, hTm, hInfo, hPos, hAlive, hPick;   // (p = pinned host memory)
    // device
    DevBuf raw, fa, faWork[2], tabs32b, tabs64b, seq2, nmask, wordContig, tabs32, tabs64, sk, br, filter, fterm, scanTmp, misc, emitMask, emitOff,
        rows0, rows0p, rows, remap, at, ovf, key1, idMeta, alive, idTm, idInfo, pos32, pick, candList;
    K k{};
    uint64_t nslots = 0, nids = 0, nrec = 0, npick = 0;
    uint64_t launches = 0;
    bool used_filter = false;
    pfb_timing tm{};
    // scratch of one run (shared by the phases)
    struct Run {
        bool piped = false, partial = false, use_filter1 = false;
        uint32_t W = 0, W1 = 0, N1 = 0, NB = 0;
        uint64_t W_alloc = 0, windows1 = 0;
        int n_out = 0;
        std::vector<size_t> cuts;               // ends (exclusive) of the batches of genomes 2..G
        bool packed_aside = false;              // ... and whether they are being packed on the side stream
        bool rows_zeroed = false;               // the rows of genomes 2..G have been zeroed on the side stream (by capacity)
    } run;
};

struct pfb_multi {
    std::vector<pfb_ctx *> ctx;
    std::vector<std::pair<int, uint32_t>> where;   // global genome index -> (context, local index); genome 0: (0, 0) and everywhere
    std::string err;
};

static std::string g_create_err;

namespace {

void trace_mark(pfb_ctx *ctx, cudaStream_t st, const std::string &label) {
    if (!ctx->tracing) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ctx->trace.emplace_back(label, e);
}

void trace_dump(pfb_ctx *ctx) {
    if (!ctx->tracing) return;
    for (auto &t : ctx->trace) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev[ctx->run.piped ? 0 : 2], t.second) == cudaSuccess)
            fprintf(stderr, "[pfb trace r%d] %8.3f ms  %s\n", ctx->rank, ms, t.first.c_str());
        cudaEventDestroy(t.second);
    }
    cudaGetLastError();
    ctx->trace.clear();
}

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

// grow a buffer keeping its first `keep` bytes
int ensure_keep(pfb_ctx *ctx, DevBuf &b, size_t bytes, size_t keep) {
    if (bytes <= b.cap && b.p) return PFB_OK;
    DevBuf old = b;
    b.p = nullptr; b.cap = 0;
    int rc = ensure(ctx, b, bytes);
    if (rc != PFB_OK) return rc;
    if (old.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (keep) CK(cudaMemcpy(b.p, old.p, std::min(keep, old.cap), cudaMemcpyDeviceToDevice));
        CK(cudaFree(old.p));
    }
    return PFB_OK;
}

// zero the first `bytes` of up to four DevBufs in one launch (rounded up to 16: every buffer has that much slack,
// see ensure)
struct ZeroBatch {
    ZeroArgs a{};
    size_t total = 0;
    int add(pfb_ctx *ctx, DevBuf &b, size_t bytes);
    int launch(pfb_ctx *ctx, cudaStream_t st);
};

int ZeroBatch::add(pfb_ctx *ctx, DevBuf &b, size_t bytes) {
    if (!bytes) return PFB_OK;
    const size_t n16 = (bytes + 15) / 16;
    if (n16 * 16 > b.cap || a.cnt >= 4) return fail(ctx, PFB_ECUDA, "zero fill beyond the buffer");
    a.p[a.cnt] = (uint4 *)b.p; a.n16[a.cnt] = n16; a.cnt++;
    total += n16;
    return PFB_OK;
}

int ZeroBatch::launch(pfb_ctx *ctx, cudaStream_t st) {
    if (!a.cnt) return PFB_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16);
    k_zero<<<blocks, 256, 0, st>>>(a);
    ctx->launches++;
    a.cnt = 0; total = 0;
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
    if (p->table_size < 64 || p->table_size >= (1u << 31)) { why = "table_size out of range"; return PFB_EINVAL; }
    if (p->max_repeat < 1) { why = "max_repeat must be >= 1"; return PFB_EINVAL; }
    return PFB_OK;
}

// constants of oligotm() and the GC window, computed with the host libm / IEEE doubles exactly like
// the reference's C and Python code does
// fterm upload: 33*33 doubles, then (16-byte aligned) 33*33 float4 Tm window lines
constexpr size_t TMB_OFF = 33 * 33 + 1, FTERM_DOUBLES = TMB_OFF + 33 * 33 * 2;
void fill_filter_constants(const pfb_params &P, K &k, std::vector<double> &fterm) {
    // fterm holds 33*33 doubles followed by 33*33 float4 (the Tm window lines), one upload
    double dv = P.dv_conc, dntp = P.dntp_conc;
    if (dv == 0) dntp = 0;
    if (dv < dntp) dv = dntp;
    double k_mm = P.mv_conc + 120.0 * std::sqrt(dv - dntp);
    k.ln_salt = std::log(k_mm / 1000.0);
    k.ln_dna4 = std::log(P.dna_conc / 4000000000.0);
    k.ln_dna1 = std::log(P.dna_conc / 1000000000.0);
    k.dmso_term = P.dmso_conc * P.dmso_fact;
    k.min_tm = P.min_tm; k.max_tm = P.max_tm;
    fterm.assign(FTERM_DOUBLES, 0.0);
    float *tmb = reinterpret_cast<float *>(fterm.data() + TMB_OFF);
    for (int kk = 1; kk <= 32; kk++) {
        uint64_t ok = 0;
        for (int gc = 0; gc <= kk; gc++) {
            double per = (double)gc / kk * 100;                                   // Primer.py:100
            if (per >= P.min_gc && per <= P.max_gc) ok |= 1ull << gc;              // getCandidateKmers.py:303
            fterm[kk * 33 + gc] = (0.453 * gc / kk - 2.88) * P.formamide_conc;   // oligotm.c: 0.453 * GC_count / len, left to right
            // Tm >= t  <=>  dh >= 0.001 T ds - T A / 100 with T = t + 273.15 + dmso - fterm, A = salt and DNA terms
            const double A = 0.368 * (kk - 1) * k.ln_salt + 1.987 * k.ln_dna4;
            const double c = 273.15 + k.dmso_term - fterm[kk * 33 + gc];
            const double Tmin = P.min_tm + c, Tmax = P.max_tm + c;
            float *o = tmb + (size_t)(kk * 33 + gc) * 4;
            o[0] = (float)(0.001 * Tmin); o[1] = (float)(-Tmin * A / 100.0);
            o[2] = (float)(0.001 * Tmax); o[3] = (float)(-Tmax * A / 100.0);
        }
        k.gc_ok[kk] = ok;
    }
    k.gc_ok[0] = 0;
}

// where every genome's bases start in the raw arena (FASTA genomes: the file size bounds the bases)
void raw_starts(pfb_ctx *ctx) {
    ctx->gRawStart.clear();
    ctx->faStart.clear();
    uint64_t raw = 0, fa = 0;
    for (auto &g : ctx->genomes) {
        raw = (raw + 255) & ~255ull;
        fa = (fa + 255) & ~255ull;
        ctx->gRawStart.push_back(raw);
        ctx->faStart.push_back(fa);
        raw += g.fasta ? g.file_bytes : g.n_bases;
        if (g.fasta) fa += g.file_bytes;
    }
    ctx->totalRaw = raw;
    ctx->totalFa = fa;
}

// layout of the arena for the first n genomes of the host-side genome list
int build_layout(pfb_ctx *ctx, size_t n) {
    auto &G = ctx->genomes;
    ctx->cWordStart.clear(); ctx->cLen.clear(); ctx->cGenome.clear(); ctx->cVirt.clear(); ctx->cRawStart.clear();
    ctx->gWordStart.clear(); ctx->gFirstContig.clear(); ctx->gNumContigs.clear(); ctx->gVirtLen.clear();
    uint64_t words = 0;
    for (size_t g = 0; g < n; g++) {
        const uint64_t raw = ctx->gRawStart[g];
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
        if (words >= (1ull << 31)) return fail(ctx, PFB_ETOOLARGE, "more than 2^36 bases on one context");
    }
    ctx->totalWords = (uint32_t)words;
    ctx->laid_out = n;
    return PFB_OK;
}

// contig / genome tables of the current layout -> device; `second` uses the other pair of buffers so that
// kernels already launched with the first tables keep reading valid memory
int upload_tables(pfb_ctx *ctx, bool second) {
    cudaStream_t st = ctx->stream;
    DevBuf &d32 = second ? ctx->tabs32b : ctx->tabs32;
    DevBuf &d64 = second ? ctx->tabs64b : ctx->tabs64;
    std::vector<uint32_t> &t32 = second ? ctx->t32hostb : ctx->t32host;
    const size_t nC = ctx->cLen.size(), nG = ctx->laid_out;
    t32.clear();
    auto put32 = [&](const std::vector<uint32_t> &v) { size_t o = t32.size(); t32.insert(t32.end(), v.begin(), v.end()); return o; };
    size_t o_cws = put32(ctx->cWordStart), o_cl = put32(ctx->cLen), o_cg = put32(ctx->cGenome), o_cv = put32(ctx->cVirt);
    size_t o_gws = put32(ctx->gWordStart), o_gfc = put32(ctx->gFirstContig), o_gnc = put32(ctx->gNumContigs), o_gvl = put32(ctx->gVirtLen);
    int rc;
    if ((rc = ensure(ctx, d32, t32.size() * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, d64, nC * 8 + 8)) != PFB_OK) return rc;
    std::vector<uint64_t> &t64 = second ? ctx->t64hostb : ctx->t64host;
    t64 = ctx->cRawStart;
    if (!t32.empty()) CK(cudaMemcpyAsync(d32.p, t32.data(), t32.size() * 4, cudaMemcpyHostToDevice, st));
    if (nC) CK(cudaMemcpyAsync(d64.p, t64.data(), nC * 8, cudaMemcpyHostToDevice, st));
    const uint32_t *b32 = (const uint32_t *)d32.p;
    K &k = ctx->k;
    k.cWordStart = b32 + o_cws; k.cLen = b32 + o_cl; k.cGenome = b32 + o_cg; k.cVirt = b32 + o_cv;
    k.gWordStart = b32 + o_gws; k.gFirstContig = b32 + o_gfc; k.gNumContigs = b32 + o_gnc; k.gVirtLen = b32 + o_gvl;
    k.cRawStart = (const uint64_t *)d64.p;
    k.nContigs = (uint32_t)nC; k.nGenomes = (uint32_t)nG; k.totalWords = ctx->totalWords;
    k.raw = (const uint8_t *)ctx->raw.p;
    return PFB_OK;
}

// ---- FASTA ingest (pfb_fasta.cuh): the files of genomes [g0, g1) are parsed on `ps` once their copies landed
int fasta_enqueue(pfb_ctx *ctx, int b, size_t g0, size_t g1, cudaStream_t ps) {
    FaBatch &B = ctx->faBatch[b];
    B.g0 = g0; B.g1 = g1; B.pending = false;
    const uint32_t nF = (uint32_t)(g1 - g0);
    if (!nF) return PFB_OK;
    std::vector<uint64_t> &h64 = ctx->fa64host[b];
    std::vector<uint32_t> &h32 = ctx->fa32host[b];
    h64.assign(3 * (size_t)nF, 0);
    h32.assign(nF + 1, 0);
    uint64_t bytes = 0;
    uint32_t nT = 0;
    for (uint32_t f = 0; f < nF; f++) {
        const HostGenome &hg = ctx->genomes[g0 + f];
        h64[f] = ctx->faStart[g0 + f];
        h64[nF + f] = hg.file_bytes;
        h64[2 * nF + f] = ctx->gRawStart[g0 + f];
        h32[f] = nT;
        nT += (uint32_t)((hg.file_bytes + fa::TILE - 1) / fa::TILE);
        bytes += hg.file_bytes;
    }
    h32[nF] = nT;
    B.nTiles = nT;
    if (B.cap == 0) B.cap = (uint32_t)std::min<uint64_t>(1u << 26, 65536 + bytes / 256);
    // device work area
    size_t o = 0;
    auto carve = [&](size_t n) { size_t at = o; o += (n + 255) & ~(size_t)255; return at; };
    const size_t o64 = carve(3 * (size_t)nF * 8), o32 = carve((nF + 1) * 4), oFirst = carve(nF * 4);
    const size_t oTLast = carve((size_t)nT * 8), oInLast = carve((size_t)nT * 8);
    const size_t oTMark = carve((size_t)nT * 8), oInMark = carve((size_t)nT * 8);
    const size_t oTHdr = carve((size_t)(nT + 1) * 4), oHdrBase = carve((size_t)(nT + 1) * 4);
    const size_t oTKept = carve((size_t)(nT + 1) * 4), oKeptBase = carve((size_t)(nT + 1) * 4);
    int rc;
    if ((rc = ensure(ctx, ctx->faWork[b], o)) != PFB_OK) return rc;
    // host-mapped output: per-file table, then record starts, then header offsets
    const size_t hostBytes = ((size_t)(2 * nF + 1) + 2 * (size_t)B.cap) * 4;
    if (ctx->faHostCap[b] < hostBytes) {
        if (ctx->faHost[b]) cudaFreeHost(ctx->faHost[b]);
        ctx->faHost[b] = nullptr; ctx->faHostCap[b] = 0;
        CK(cudaHostAlloc(&ctx->faHost[b], hostBytes, cudaHostAllocMapped));
        ctx->faHostCap[b] = hostBytes;
    }
    uint8_t *w = (uint8_t *)ctx->faWork[b].p;
    fa::Args &A = B.args;
    A.stage = (const uint8_t *)ctx->fa.p;
    A.src = (const uint64_t *)(w + o64); A.len = A.src + nF; A.dst = A.src + 2 * nF;
    A.tile0 = (const uint32_t *)(w + o32);
    A.nFiles = nF; A.nTiles = nT;
    A.format = ctx->genomes[g0].fasta;
    A.tLast = (uint64_t *)(w + oTLast); A.inLast = (uint64_t *)(w + oInLast);
    A.tMark = (uint64_t *)(w + oTMark); A.inMark = (uint64_t *)(w + oInMark);
    A.tHdr = (uint32_t *)(w + oTHdr); A.hdrBase = (uint32_t *)(w + oHdrBase);
    A.tKept = (uint32_t *)(w + oTKept); A.keptBase = (uint32_t *)(w + oKeptBase);
    A.firstHdr = (uint32_t *)(w + oFirst);
    A.out = (uint8_t *)ctx->raw.p;
    uint32_t *hm = nullptr;
    CK(cudaHostGetDevicePointer((void **)&hm, ctx->faHost[b], 0));
    A.fileTab = hm; A.recStart = hm + (2 * nF + 1); A.recHdr = A.recStart + B.cap;
    A.cap = B.cap;
    memset(ctx->faHost[b], 0, (size_t)(2 * nF + 1) * 4);
    if (!B.done) CK(cudaEventCreateWithFlags(&B.done, cudaEventDisableTiming));
    CK(cudaMemcpyAsync(w + o64, h64.data(), h64.size() * 8, cudaMemcpyHostToDevice, ps));
    CK(cudaMemcpyAsync(w + o32, h32.data(), h32.size() * 4, cudaMemcpyHostToDevice, ps));
    CK(cudaMemsetAsync(w + oFirst, 0xFF, nF * 4, ps));
    if (nT) {
        const unsigned nb = nblk(nT, fa::WARPS);
        if (A.format == fa::FMT_FASTA) {
            fa::k_fa_mark<fa::FMT_FASTA><<<nb, fa::WARPS * 32, 0, ps>>>(A);
            fa::k_fa_scan<<<1, 1024, 0, ps>>>(A.tHdr, A.hdrBase, A.tLast, A.inLast, nullptr, nullptr, nT);
            fa::k_fa_sweep<fa::FMT_FASTA, false><<<nb, fa::WARPS * 32, 0, ps>>>(A);
            fa::k_fa_scan<<<1, 1024, 0, ps>>>(A.tKept, A.keptBase, nullptr, nullptr, nullptr, nullptr, nT);
            fa::k_fa_sweep<fa::FMT_FASTA, true><<<nb, fa::WARPS * 32, 0, ps>>>(A);
        } else {
            fa::k_fa_mark<fa::FMT_GENBANK><<<nb, fa::WARPS * 32, 0, ps>>>(A);
            fa::k_fa_scan<<<1, 1024, 0, ps>>>(A.tHdr, A.hdrBase, A.tLast, A.inLast, A.tMark, A.inMark, nT);
            fa::k_fa_sweep<fa::FMT_GENBANK, false><<<nb, fa::WARPS * 32, 0, ps>>>(A);
            fa::k_fa_scan<<<1, 1024, 0, ps>>>(A.tKept, A.keptBase, nullptr, nullptr, nullptr, nullptr, nT);
            fa::k_fa_sweep<fa::FMT_GENBANK, true><<<nb, fa::WARPS * 32, 0, ps>>>(A);
        }
        ctx->fa_launches += 5;
    }
    CK(cudaEventRecord(B.done, ps));
    B.pending = true;
    return PFB_OK;
}

// wait for a batch and turn its record table into contig offsets of the genomes
int fasta_collect(pfb_ctx *ctx, int b) {
    FaBatch &B = ctx->faBatch[b];
    if (!B.pending) return PFB_OK;
    B.pending = false;
    CK(cudaEventSynchronize(B.done));
    const uint32_t nF = B.args.nFiles;
    const uint32_t *tab = (const uint32_t *)ctx->faHost[b];
    const uint32_t total = tab[2 * nF];
    if (total > B.cap) {
        // more records than the table holds: grow it and repeat the gather pass (the scans are still valid)
        B.cap = total + total / 4 + 16;
        const size_t hostBytes = ((size_t)(2 * nF + 1) + 2 * (size_t)B.cap) * 4;
        cudaFreeHost(ctx->faHost[b]);
        ctx->faHost[b] = nullptr; ctx->faHostCap[b] = 0;
        CK(cudaHostAlloc(&ctx->faHost[b], hostBytes, cudaHostAllocMapped));
        ctx->faHostCap[b] = hostBytes;
        uint32_t *hm = nullptr;
        CK(cudaHostGetDevicePointer((void **)&hm, ctx->faHost[b], 0));
        B.args.fileTab = hm; B.args.recStart = hm + (2 * nF + 1); B.args.recHdr = B.args.recStart + B.cap;
        B.args.cap = B.cap;
        if (B.args.format == fa::FMT_FASTA) fa::k_fa_sweep<fa::FMT_FASTA, true><<<nblk(B.nTiles, fa::WARPS), fa::WARPS * 32, 0, ctx->parseStream>>>(B.args);
        else fa::k_fa_sweep<fa::FMT_GENBANK, true><<<nblk(B.nTiles, fa::WARPS), fa::WARPS * 32, 0, ctx->parseStream>>>(B.args);
        ctx->fa_launches++;
        CK(cudaStreamSynchronize(ctx->parseStream));
        tab = (const uint32_t *)ctx->faHost[b];
    }
    const uint32_t *recStart = tab + (2 * nF + 1), *recHdr = recStart + B.cap;
    for (uint32_t f = 0; f < nF; f++) {
        HostGenome &hg = ctx->genomes[B.g0 + f];
        const uint32_t r0 = tab[2 * f], kept = tab[2 * f + 1];
        const uint32_t r1 = f + 1 < nF ? tab[2 * (f + 1)] : total;
        if (r1 <= r0) return fail(ctx, PFB_EINVAL, "genome " + std::to_string(B.g0 + f) + ": no record (no line starts with '>' / LOCUS)");
        hg.off.clear(); hg.hdr_off.clear();
        for (uint32_t r = r0; r < r1; r++) { hg.off.push_back(recStart[r]); hg.hdr_off.push_back(recHdr[r]); }
        hg.off.push_back(kept);
        hg.n_bases = kept;
    }
    return PFB_OK;
}

// device-wide exclusive scan helper (out may alias in)
template <int MODE, int S>
int device_exscan(pfb_ctx *ctx, const void *in, uint32_t *out, uint32_t n, uint32_t *total_dev, uint32_t *total_host) {
    cudaStream_t st = ctx->stream;
    uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    int rc = ensure(ctx, ctx->scanTmp, (size_t)(nb + 1) * 4);
    if (rc != PFB_OK) return rc;
    uint32_t *bs = (uint32_t *)ctx->scanTmp.p;
    k_scan_p1<MODE, S><<<nb, 256, 0, st>>>(in, n, bs);
    k_scan_p3<MODE, S><<<nb, 256, 0, st>>>(in, out, n, bs, total_dev, total_host);
    ctx->launches += 2;
    return PFB_OK;
}

// NCCL entry points, bound once
nccl::Api *nccl_api(std::string &why) {
    static nccl::Api api;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (!api.handle) {
            err = std::string("NCCL is needed for more than one GPU and could not be loaded: ") + dlerror();
        } else {
            bool ok = true;
            auto sym = [&](const char *name) { void *f = dlsym(api.handle, name); if (!f) { ok = false; err = std::string("NCCL lacks ") + name; } return f; };
            api.GetUniqueId = (int (*)(nccl::UniqueId *))sym("ncclGetUniqueId");
            api.CommInitRank = (int (*)(nccl::Comm **, int, nccl::UniqueId, int))sym("ncclCommInitRank");
            api.CommInitAll = (int (*)(nccl::Comm **, int, const int *))sym("ncclCommInitAll");
            api.CommDestroy = (int (*)(nccl::Comm *))sym("ncclCommDestroy");
            api.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl::Comm *, cudaStream_t))sym("ncclAllReduce");
            api.AllGather = (int (*)(const void *, void *, size_t, int, nccl::Comm *, cudaStream_t))sym("ncclAllGather");
            api.GroupStart = (int (*)())sym("ncclGroupStart");
            api.GroupEnd = (int (*)())sym("ncclGroupEnd");
            api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
            if (!ok) { dlclose(api.handle); api.handle = nullptr; }
        }
    }
    if (!api.handle) { why = err; return nullptr; }
    return &api;
}

}  // namespace

// ==========================================================================================
extern "C" {

// the peer arena: wait (bounded) until no other rank can still be reading it, unmap the others', free
static void x_release(pfb_ctx *ctx) {
    if (!ctx->xa.p) return;
    cudaSetDevice(ctx->device);
    if (ctx->peer_now && ctx->epoch && ctx->nRanks > 1) {
        k_xwait<<<1, 32, 0, ctx->stream>>>(ctx->k, XP_DONE, ctx->epoch);
        cudaStreamSynchronize(ctx->stream);
    }
    for (int r = 0; r < 16; r++) {
        if (ctx->xOpened[r] && ctx->xPeer[r]) cudaIpcCloseMemHandle(ctx->xPeer[r]);
        ctx->xOpened[r] = false; ctx->xPeer[r] = nullptr;
    }
    release(ctx->xa);
    cudaGetLastError();
    ctx->peer_now = false; ctx->epoch = 0;
    ctx->xCaps[0] = ctx->xCaps[1] = ctx->xCaps[2] = 0;
}

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
    { const char *tr = getenv("PFB200_TRACE"); ctx->tracing = tr && tr[0] == '1'; }
    { const char *aw = getenv("PFB200_AT_WIDE"); if (aw && atoi(aw) > 0) ctx->at_wide = std::min<uint32_t>(AT_SLOW, (uint32_t)atoi(aw)); }
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->parseStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->d2hStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evSide, cudaEventDisableTiming);
    { const char *sg = getenv("PFB200_SHARD_G1"); if (sg && sg[0] == '0') ctx->shard_g1 = 0; }
    { const char *cs = getenv("PFB200_CAP_SHARE"); if (cs && atof(cs) > 0) ctx->cap_share = atof(cs); }
    { const char *xc = getenv("PFB200_XCHG"); if (xc && (xc[0] == 'n' || xc[0] == 'N')) ctx->xchg_pref = 0; }
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_scan_filtered, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCANF_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_scan_filtered2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN2_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_scan_filtered2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN2_SMEM);
    { const char *sk = getenv("PFB200_SCAN_KERNEL"); if (sk && sk[0] == '1') ctx->scan_kernel = 1; }
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete ctx; return fail(nullptr, PFB_ECUDA, m); }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    for (auto &pair : ctx->evX) for (auto &ev : pair) cudaEventCreate(&ev);
    {
        void *h = nullptr, *d = nullptr;
        if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
            std::string m = "cudaHostAlloc (mapped) failed"; pfb_destroy(ctx); return fail(nullptr, PFB_ECUDA, m);
        }
        memset(h, 0, 64);
        ctx->countsHost = (volatile uint32_t *)h; ctx->countsDev = (uint32_t *)d;
    }
    *out = ctx;
    return PFB_OK;
}

void pfb_destroy(pfb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d2hStream) cudaStreamSynchronize(ctx->d2hStream);
    x_release(ctx);
    if (ctx->comm && ctx->comm_owned) { std::string why; if (nccl::Api *a = nccl_api(why)) a->CommDestroy(ctx->comm); }
    ctx->comm = nullptr;
    DevBuf *all[] = {&ctx->raw, &ctx->seq2, &ctx->nmask, &ctx->wordContig, &ctx->tabs32, &ctx->tabs64, &ctx->sk, &ctx->br,
                     &ctx->filter, &ctx->fterm, &ctx->scanTmp, &ctx->misc, &ctx->emitMask, &ctx->emitOff,
                     &ctx->rows0, &ctx->rows0p, &ctx->rows, &ctx->remap, &ctx->at, &ctx->ovf, &ctx->key1, &ctx->idMeta, &ctx->alive,
                     &ctx->idTm, &ctx->idInfo, &ctx->pos32, &ctx->pick, &ctx->candList};
    for (auto *b : all) release(*b);
    DevBuf *hosts[] = {&ctx->hWord, &ctx->hTm, &ctx->hInfo, &ctx->hPos, &ctx->hAlive, &ctx->hPick};
    for (auto *b : hosts) { if (b->p) cudaFreeHost(b->p); b->p = nullptr; b->cap = 0; }
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto &pair : ctx->evX) for (auto &ev : pair) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->evCopy) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->evOut) if (ev) cudaEventDestroy(ev);
    if (ctx->d2hStream) cudaStreamDestroy(ctx->d2hStream);
    if (ctx->sideStream) { cudaStreamSynchronize(ctx->sideStream); cudaStreamDestroy(ctx->sideStream); }
    if (ctx->evSide) cudaEventDestroy(ctx->evSide);
    for (auto &ev : ctx->evEnc) if (ev) cudaEventDestroy(ev);
    if (ctx->countsHost) cudaFreeHost((void *)ctx->countsHost);
    if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
    if (ctx->parseStream) { cudaStreamSynchronize(ctx->parseStream); cudaStreamDestroy(ctx->parseStream); }
    for (int b = 0; b < 2; b++) {
        release(ctx->faWork[b]);
        if (ctx->faHost[b]) cudaFreeHost(ctx->faHost[b]);
        if (ctx->faBatch[b].done) cudaEventDestroy(ctx->faBatch[b].done);
    }
    release(ctx->fa); release(ctx->tabs32b); release(ctx->tabs64b);
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
    ctx->uploaded = false; ctx->computed = false; ctx->host_valid = false;
    return PFB_OK;
}

static int add_file(pfb_ctx *ctx, const uint8_t *file_bytes, uint64_t n_bytes, int format) {
    if (!ctx) return PFB_EINVAL;
    if (!file_bytes || n_bytes == 0) return fail(ctx, PFB_EINVAL, "empty file");
    if (n_bytes >= (1ull << 31)) return fail(ctx, PFB_ETOOLARGE, "file of 2 GiB or more");
    HostGenome g;
    g.bases = file_bytes; g.n_bases = 0; g.fasta = format; g.file_bytes = n_bytes;
    ctx->genomes.push_back(std::move(g));
    ctx->uploaded = false; ctx->computed = false; ctx->host_valid = false;
    return PFB_OK;
}

int pfb_add_fasta(pfb_ctx *ctx, const uint8_t *file_bytes, uint64_t n_bytes) { return add_file(ctx, file_bytes, n_bytes, fa::FMT_FASTA); }

int pfb_add_genbank(pfb_ctx *ctx, const uint8_t *file_bytes, uint64_t n_bytes) { return add_file(ctx, file_bytes, n_bytes, fa::FMT_GENBANK); }

int pfb_genome_info(pfb_ctx *ctx, uint32_t genome_idx, uint32_t *n_contigs, uint64_t *n_bases) {
    if (!ctx) return PFB_EINVAL;
    if (genome_idx >= ctx->genomes.size()) return fail(ctx, PFB_EINVAL, "genome index out of range");
    const HostGenome &g = ctx->genomes[genome_idx];
    if (g.off.empty()) return fail(ctx, PFB_ESTATE, "the genome has not been parsed yet (pfb_upload / pfb_run)");
    if (n_contigs) *n_contigs = (uint32_t)g.off.size() - 1;
    if (n_bases) *n_bases = g.n_bases;
    return PFB_OK;
}

int pfb_contig_table(pfb_ctx *ctx, uint32_t genome_idx, uint64_t *header_off, uint64_t *length, uint32_t cap) {
    if (!ctx) return PFB_EINVAL;
    if (genome_idx >= ctx->genomes.size()) return fail(ctx, PFB_EINVAL, "genome index out of range");
    const HostGenome &g = ctx->genomes[genome_idx];
    if (g.off.empty()) return fail(ctx, PFB_ESTATE, "the genome has not been parsed yet (pfb_upload / pfb_run)");
    const uint32_t nc = (uint32_t)g.off.size() - 1;
    if (cap < nc) return fail(ctx, PFB_EINVAL, "contig table too small");
    for (uint32_t c = 0; c < nc; c++) {
        if (header_off) header_off[c] = g.fasta ? g.hdr_off[c] : 0;
        if (length) length[c] = g.off[c + 1] - g.off[c];
    }
    return PFB_OK;
}

int pfb_clear_genomes(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    ctx->genomes.clear();
    ctx->uploaded = false; ctx->computed = false; ctx->host_valid = false;
    return PFB_OK;
}

int pfb_sync(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->d2hStream));
    return PFB_OK;
}

int pfb_stream(pfb_ctx *ctx, void **stream_out) {
    if (!ctx || !stream_out) return PFB_EINVAL;
    *stream_out = (void *)ctx->stream;
    return PFB_OK;
}

int pfb_set_scan_mode(pfb_ctx *ctx, int mode) {
    if (!ctx) return PFB_EINVAL;
    if (mode < 0 || mode > 2) return fail(ctx, PFB_EINVAL, "scan mode must be 0 (auto), 1 (L2 probe) or 2 (shared-memory filter)");
    ctx->scan_mode = mode;
    return PFB_OK;
}

static int upload_impl(pfb_ctx *ctx, bool async) {
    if (!ctx) return PFB_EINVAL;
    if (ctx->genomes.empty()) return fail(ctx, PFB_ESTATE, "no genome added");
    CK(cudaSetDevice(ctx->device));
    const size_t nG = ctx->genomes.size();
    size_t nFasta = 0;
    for (auto &g : ctx->genomes) nFasta += g.fasta == ctx->genomes[0].fasta && g.fasta ? 1 : 0;
    for (auto &g : ctx->genomes)
        if (g.fasta != ctx->genomes[0].fasta)
            return fail(ctx, PFB_EINVAL, "pfb_add_genome, pfb_add_fasta and pfb_add_genbank cannot be mixed on one context");
    ctx->any_fasta = nFasta != 0;
    raw_starts(ctx);
    int rc;
    cudaStream_t st = ctx->stream;
    cudaStream_t cs = async ? ctx->copyStream : ctx->stream;
    if ((rc = ensure(ctx, ctx->raw, ctx->totalRaw + 64)) != PFB_OK) return rc;
    if (ctx->any_fasta && (rc = ensure(ctx, ctx->fa, ctx->totalFa + 256)) != PFB_OK) return rc;
    while (ctx->evCopy.size() < nG) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->evCopy.push_back(e);
    }
    ctx->fa_launches = 0;
    if (!ctx->any_fasta) {   // the (small) table copies go first: behind the bases they would wait for a 5 MB chunk
        if ((rc = build_layout(ctx, nG)) != PFB_OK) return rc;
        if ((rc = upload_tables(ctx, false)) != PFB_OK) return rc;
    }
    CK(cudaEventRecord(ctx->ev[0], cs));
    for (size_t g = 0; g < nG; g++) {
        auto &hg = ctx->genomes[g];
        if (hg.fasta) {
            CK(cudaMemcpyAsync((uint8_t *)ctx->fa.p + ctx->faStart[g], hg.bases, hg.file_bytes, cudaMemcpyHostToDevice, cs));
        } else if (hg.n_bases) {
            CK(cudaMemcpyAsync((uint8_t *)ctx->raw.p + ctx->gRawStart[g], hg.bases, hg.n_bases, cudaMemcpyHostToDevice, cs));
        }
        if (async) CK(cudaEventRecord(ctx->evCopy[g], cs));
        trace_mark(ctx, cs, "copy landed: genome " + std::to_string(g));
    }
    CK(cudaEventRecord(ctx->ev[1], cs));
    ctx->pipelined = async;
    if (!ctx->any_fasta) {
    } else if (!async) {
        // every file is parsed on the device before the layout is made
        if ((rc = fasta_enqueue(ctx, 0, 0, nG, st)) != PFB_OK) return rc;
        if ((rc = fasta_collect(ctx, 0)) != PFB_OK) return rc;
        if ((rc = build_layout(ctx, nG)) != PFB_OK) return rc;
        if ((rc = upload_tables(ctx, false)) != PFB_OK) return rc;
    } else {
        // genome 1's file is parsed as soon as it has landed and the key table is built from it while the
        // other files are still in flight; they are parsed behind their copies and join the layout just
        // before they are scanned (stage0 -> fasta_finish)
        CK(cudaStreamWaitEvent(ctx->parseStream, ctx->evCopy[0], 0));
        if ((rc = fasta_enqueue(ctx, 0, 0, 1, ctx->parseStream)) != PFB_OK) return rc;
        if (nG > 1) {
            CK(cudaStreamWaitEvent(ctx->parseStream, ctx->evCopy[nG - 1], 0));
            if ((rc = fasta_enqueue(ctx, 1, 1, nG, ctx->parseStream)) != PFB_OK) return rc;
        }
        if ((rc = fasta_collect(ctx, 0)) != PFB_OK) return rc;
        if ((rc = build_layout(ctx, 1)) != PFB_OK) return rc;
        if ((rc = upload_tables(ctx, false)) != PFB_OK) return rc;
    }
    if (!async) {
        CK(cudaStreamSynchronize(st));   // host genome buffers are borrowed only until here
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        ctx->tm.ms_upload = ms;
    }
    ctx->uploaded = true; ctx->computed = false; ctx->host_valid = false;
    return PFB_OK;
}

// FASTA, pipelined: the files of genomes 2..G have been parsed by now (or are about to be): complete the layout
static int fasta_finish(pfb_ctx *ctx) {
    int rc;
    const size_t nG = ctx->genomes.size();
    if (ctx->laid_out == nG) return PFB_OK;
    if ((rc = fasta_collect(ctx, 1)) != PFB_OK) return rc;
    if ((rc = build_layout(ctx, nG)) != PFB_OK) return rc;
    return upload_tables(ctx, true);
}

// copies may be in flight when the upload fails half way (e.g. a file without records): the host buffers are
// borrowed until they are done, error or not
static int upload_checked(pfb_ctx *ctx, bool async) {
    int rc = upload_impl(ctx, async);
    if (rc != PFB_OK && ctx) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copyStream);
        cudaStreamSynchronize(ctx->parseStream);
        cudaGetLastError();
        ctx->pipelined = false;
    }
    return rc;
}

int pfb_upload(pfb_ctx *ctx) { return upload_checked(ctx, false); }

int pfb_upload_async(pfb_ctx *ctx) { return upload_checked(ctx, true); }

int pfb_upload_wait(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copyStream));
    if (ctx->pipelined) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->tm.ms_upload = ms; else cudaGetLastError();
    }
    ctx->pipelined = false;
    return PFB_OK;
}

// ------------------------------------------------------------------------------------------
// One run = five phases of kernels separated by (at most) four collectives.  No phase waits for the device: the
// buffers are sized by capacities (pick_capacities), the kernels read the actual counts from device memory, and
// a count that did not fit raises a flag that the host finds at the very end -- it then grows the capacities and
// repeats the run.  With one context the collectives are no-ops; with several contexts in one process
// (pfb_multi) one host thread walks all of them through each phase and issues the collectives as one NCCL group;
// with one process per GPU every rank does the same on its own context.
//
//   A  zero fills, genome 1 packed, genome 1's candidate windows (its share of them when the ranks share the work)
//      -- all-gather of the candidate lists --
//   B  candidates -> table, key bitmap + ranks, genome 1's own pass (its share)
//      -- SUM all-reduce of genome 1's packed rows --
//   C  genome 1's verdicts, ids in dict order, key table of the other genomes, the per-id records; then the
//      other genomes in batches: pack, scan, uniqueness rules, places (results leave for the host as they are made)
//      -- MIN all-reduce of the alive bytes --
//   D  per-position pick
//      -- MAX all-reduce of the pick bytes --
//   E  counts, last result copies

static void pick_capacities(pfb_ctx *ctx, uint64_t windows1) {
    const pfb_params &P = ctx->prm;
    const int nk = P.k_max - P.k_min + 1;
    auto round256 = [](uint64_t x) { return (uint32_t)std::min<uint64_t>((x + 255) & ~255ull, 0xFFFFFF00ull); };
    if (ctx->caps_for_prefilter != P.prefilter || ctx->caps_for_windows != windows1) {
        // nothing known yet about these inputs: a generous share of genome 1's windows (every key is a candidate
        // window of genome 1; with the prefilter ~9 % of the windows of a random 50 % GC genome are candidates)
        const double share = ctx->cap_share > 0 ? ctx->cap_share : (P.prefilter ? 0.14 : 0.80);
        const uint64_t est = std::min<uint64_t>((uint64_t)(share * (double)windows1) + (ctx->cap_share > 0 ? 256 : 65536), (uint64_t)nk * P.table_size);
        ctx->capSlots = round256(est);
        ctx->capIds = ctx->capSlots;
        ctx->caps_for_prefilter = P.prefilter;
        ctx->caps_for_windows = windows1;
        ctx->lastSlots = ctx->lastIds = 0;
    } else if (ctx->lastSlots) {
        // same shape of input as the last completed run: its counts + 6 % (a bench loop, a service that sees similar
        // genomes); a run that needs more finds out and repeats itself
        ctx->capSlots = round256(ctx->lastSlots + ctx->lastSlots / 16 + 1024);
        ctx->capIds = round256(ctx->lastIds + ctx->lastIds / 16 + 1024);
    }
    const double share = ctx->cap_share > 0 ? ctx->cap_share : (P.prefilter ? 0.14 : 0.80);
    ctx->capCand = std::max(ctx->capCandMin, round256((uint64_t)(share * (double)windows1 / std::max(1, ctx->eff_ranks())) + (ctx->cap_share > 0 ? 0 : 4096)));
}

static int ensure_host(pfb_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return PFB_OK;
    CK(cudaStreamSynchronize(ctx->d2hStream));
    if (b.p) CK(cudaFreeHost(b.p));
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    CK(cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
    b.cap = want;
    return PFB_OK;
}

static cudaEvent_t out_event(pfb_ctx *ctx) {
    if ((size_t)ctx->run.n_out >= ctx->evOut.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ctx->evOut.push_back(e);
    }
    return ctx->evOut[ctx->run.n_out++];
}

// device -> host copy of a result piece behind everything queued on the compute stream so far
static int copy_out(pfb_ctx *ctx, void *dst, const void *src, size_t bytes, const char *what = "") {
    if (!bytes) return PFB_OK;
    cudaEvent_t e = out_event(ctx);
    if (!e) return fail(ctx, PFB_ECUDA, "cudaEventCreate failed");
    CK(cudaEventRecord(e, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->d2hStream, e, 0));
    if (ctx->tracing) trace_mark(ctx, ctx->d2hStream, std::string("result copy starts: ") + what);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->d2hStream));
    if (ctx->tracing) trace_mark(ctx, ctx->d2hStream, std::string("result copy landed: ") + what + " (" + std::to_string(bytes >> 10) + " KiB)");
    return PFB_OK;
}

static int host_buffers(pfb_ctx *ctx) {
    const size_t nG = ctx->genomes.size(), CI = ctx->capIds;
    int rc;
    if ((rc = ensure_host(ctx, ctx->hWord, CI * 8)) != PFB_OK) return rc;
    if ((rc = ensure_host(ctx, ctx->hTm, CI * 8)) != PFB_OK) return rc;
    if ((rc = ensure_host(ctx, ctx->hInfo, CI * 2)) != PFB_OK) return rc;
    if ((rc = ensure_host(ctx, ctx->hPos, CI * nG * 4)) != PFB_OK) return rc;
    if ((rc = ensure_host(ctx, ctx->hAlive, CI)) != PFB_OK) return rc;
    if ((rc = ensure_host(ctx, ctx->hPick, CI)) != PFB_OK) return rc;
    ctx->hostPitch = CI;
    return PFB_OK;
}

// pipelined run: the batches of genomes 2..G that finish earliest under a simple model -- copies land in order at
// ~45 GB/s from t = 0, genome 1's pipeline takes a fixed part plus a part per window, a batch costs a fixed part
// plus its windows at the hot kernel's rate and starts when the previous one is done and its own last copy has
// landed.  Returns the end (exclusive) of every batch.  Only the SPLIT depends on the model, never a result.
static std::vector<size_t> plan_batches(const pfb_ctx *ctx) {
    const size_t nG = ctx->genomes.size();
    const double nk = ctx->k.nk;
    std::vector<double> land(nG + 1, 0.0), bases(nG + 1, 0.0);
    double acc = 0;
    for (size_t g = 0; g < nG; g++) {
        const HostGenome &hg = ctx->genomes[g];
        acc += (double)(hg.fasta ? hg.file_bytes : hg.n_bases);
        land[g] = acc / 45e9;
        bases[g + 1] = bases[g] + (double)(hg.fasta ? hg.file_bytes : hg.n_bases);
    }
    const double n1 = (double)(ctx->genomes[0].fasta ? ctx->genomes[0].file_bytes : ctx->genomes[0].n_bases);
    const double t0 = land[0] + 250e-6 + n1 * nk * 12e-12;
    const double fixed = 90e-6, per_window = 1.0 / 330e9;
    const size_t n = nG - 1;                      // genomes 1..nG-1
    // coarse grid when there are many genomes (the plan is O(n^2))
    const size_t step = std::max<size_t>(1, n / 256);
    std::vector<size_t> pts;                      // candidate cut points (exclusive ends), last = nG
    for (size_t g = 1 + step; g < nG; g += step) pts.push_back(g);
    pts.push_back(nG);
    std::vector<double> best(pts.size(), 0.0);
    std::vector<int> from(pts.size(), -1);
    for (size_t i = 0; i < pts.size(); i++) {
        best[i] = 1e30;
        for (int j = -1; j < (int)i; j++) {
            const size_t a = j < 0 ? 1 : pts[j], b = pts[i];
            const double start = std::max(j < 0 ? t0 : best[j], land[b - 1]);
            const double t = start + fixed + (bases[b] - bases[a]) * nk * per_window;
            if (t < best[i]) { best[i] = t; from[i] = j; }
        }
    }
    std::vector<size_t> cuts;
    for (int i = (int)pts.size() - 1; i >= 0; i = from[i]) cuts.push_back(pts[i]);
    std::reverse(cuts.begin(), cuts.end());
    return cuts;
}

// genomes 2..G are packed on the side stream, beside genome 1's pipeline (a string of small kernels that leaves
// most of the GPU idle): the resident run in one launch, the pipelined run batch by batch behind the H2D copies
static int pack_aside(pfb_ctx *ctx) {
    pfb_ctx::Run &R = ctx->run;
    const size_t nG = ctx->genomes.size();
    K &k = ctx->k;
    if (nG < 2 || R.packed_aside) return PFB_OK;
    R.cuts.clear();
    if (R.piped) R.cuts = plan_batches(ctx); else R.cuts.push_back(nG);
    while (ctx->evEnc.size() < R.cuts.size()) {
        cudaEvent_t e = nullptr;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->evEnc.push_back(e);
    }
    CK(cudaEventRecord(ctx->evSide, ctx->stream));               // (tables uploaded, previous run's readers done)
    CK(cudaStreamWaitEvent(ctx->sideStream, ctx->evSide, 0));
    // the rows of genomes 2..G are zeroed here as well, beside genome 1's pipeline: by capacity (the id count is not
    // known yet), but off the path that the scan waits for
    CK(cudaMemsetAsync(k.rows, 0, (size_t)(nG - 1) * ctx->capIds * 8, ctx->sideStream));
    R.rows_zeroed = true;
    size_t g = 1;
    for (size_t b = 0; b < R.cuts.size(); b++) {
        const size_t g2 = R.cuts[b];
        const uint32_t wa = ctx->gWordStart[g], wb = g2 < nG ? ctx->gWordStart[g2] : R.W;
        if (R.piped) CK(cudaStreamWaitEvent(ctx->sideStream, ctx->evCopy[g2 - 1], 0));
        if (wb > wa) { k_encode<<<nblk(wb - wa, 256), 256, 0, ctx->sideStream>>>(k, wa, wb); ctx->launches++; }
        CK(cudaEventRecord(ctx->evEnc[b], ctx->sideStream));
        g = g2;
    }
    R.packed_aside = true;
    return PFB_OK;
}

static void x_signal(pfb_ctx *ctx, int phase);
static void x_wait(pfb_ctx *ctx, int phase);
static void x_done(pfb_ctx *ctx, int phase);

static int phaseA(pfb_ctx *ctx, bool piped) {
    CK(cudaSetDevice(ctx->device));
    const pfb_params &P = ctx->prm;
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    int rc;
    const int nk = P.k_max - P.k_min + 1;
    k.kmin = P.k_min; k.kmax = P.k_max; k.nk = nk;
    k.P = P.table_size; k.PW = (P.table_size + 31) / 32;
    k.PW16 = (P.table_size + 15) / 16;
    k.barrett = (uint64_t)((((unsigned __int128)1) << 64) / P.table_size);
    k.fscale = (uint64_t)FILTER_WORDS >= P.table_size
                   ? 0xFFFFFFFFu
                   : (uint32_t)((((unsigned __int128)FILTER_WORDS) << 32) / P.table_size);
    k.max_repeat = P.max_repeat; k.prefilter = P.prefilter;
    std::vector<double> fterm;
    fill_filter_constants(P, k, fterm);
    pfb_ctx::Run &R = ctx->run;
    R = pfb_ctx::Run();
    R.piped = piped;
    R.W = ctx->totalWords;
    R.W1 = ctx->laid_out > 1 ? ctx->gWordStart[1] : R.W;                       // genome 1 = words [0, W1)
    R.N1 = R.W1 * 32u;                                                         // padded end positions of genome 1
    R.NB = (uint32_t)nk * k.PW;                                                // words of the concatenated bitmaps
    ctx->launches = 0;
    // FASTA genomes still being parsed (pipelined run): the layout covers genome 1 only, the word arrays are
    // sized by an estimate of the rest (re-checked in fasta_finish)
    R.partial = ctx->laid_out < ctx->genomes.size();
    R.W_alloc = R.W;
    if (R.partial)
        for (size_t g = ctx->laid_out; g < ctx->genomes.size(); g++)
            R.W_alloc += ctx->genomes[g].file_bytes / 32 + ctx->genomes[g].file_bytes / 1024 + 4096;
    uint64_t windows1 = 0;
    for (uint32_t c = 0; c < ctx->gNumContigs[0]; c++)
        for (int kk = P.k_min; kk <= P.k_max; kk++)
            if (ctx->cLen[c] >= (uint32_t)kk) windows1 += ctx->cLen[c] - kk + 1;
    pick_capacities(ctx, windows1);
    R.windows1 = windows1;
    const size_t nG = ctx->genomes.size();
    const size_t CS = ctx->capSlots, CI = ctx->capIds;
    const size_t NA = (size_t)nk * k.PW16;
    const size_t ovfCap = std::min<size_t>(CI / 3 + (CI > ctx->at_wide ? CI - ctx->at_wide : 0), NA) + 16;
    if (ovfCap >= (1u << 24)) return fail(ctx, PFB_ETOOLARGE, "too many keys share 16-bin words for the inline id table");
    // genome 1's own pipeline is shared out when there are ranks to share it with and the candidate lists are short
    const int nR = ctx->eff_ranks();
    ctx->sharded_now = ((ctx->comm && ctx->nRanks > 1) || ctx->emulate_ranks > 1) && ctx->shard_g1 && P.prefilter &&
                       P.table_size <= (1u << 27) && nk <= 32 && ctx->scan_mode != 2;
    // buffers
    if ((rc = ensure(ctx, ctx->seq2, (size_t)(R.W_alloc + 2) * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->nmask, (size_t)(R.W_alloc + 1) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->wordContig, (size_t)(R.W_alloc + 1) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->sk, (size_t)R.NB * 2 * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->br, (size_t)R.NB * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->filter, (size_t)nk * FILTER_BYTES)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->fterm, FTERM_DOUBLES * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->misc, 4096)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->emitMask, (size_t)R.W1 * 32 * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->emitOff, (size_t)R.W1 * 32 * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->rows0, CS * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->remap, CS * 4)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->rows, CI * (nG > 1 ? nG - 1 : 1) * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->key1, CI * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->idMeta, CI * 4)) != PFB_OK) return rc;
    // more than one rank: what the ranks exchange lives in one arena that the others map (unless NCCL is asked for)
    const bool want_peer = ctx->comm && ctx->nRanks > 1 && ctx->nRanks <= 16 && ctx->xchg_pref == 1 && ctx->emulate_ranks <= 1;
    auto al256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    if (want_peer) {
        k.xCand = X_FLAGS_BYTES;
        k.xRows0p = (uint32_t)(k.xCand + al256((size_t)(ctx->capCand + 4) * 4));
        k.xRows0s = (uint32_t)(k.xRows0p + al256(CS * 4));
        k.xAlive = (uint32_t)(k.xRows0s + al256(CS * 4));
        k.xPick = (uint32_t)(k.xAlive + al256(CI));
        const size_t total = (size_t)k.xPick + al256(CI);
        if (total >= (1ull << 32)) return fail(ctx, PFB_ETOOLARGE, "exchange arena of 4 GiB or more");
        if (!ctx->xa.p || ctx->xCaps[0] != ctx->capCand || ctx->xCaps[1] != CS || ctx->xCaps[2] != CI) {
            x_release(ctx);                                    // (waits until no peer can still be reading the old one)
            if ((rc = ensure(ctx, ctx->xa, total)) != PFB_OK) return rc;
            CK(cudaMemsetAsync(ctx->xa.p, 0, X_FLAGS_BYTES, st));
            CK(cudaStreamSynchronize(st));
            ctx->xCaps[0] = ctx->capCand; ctx->xCaps[1] = (uint32_t)CS; ctx->xCaps[2] = (uint32_t)CI;
            ctx->x_dirty = true;
            ctx->epoch = 0;
        }
        ctx->peer_now = true;
        ctx->epoch++;
    } else {
        if (ctx->xa.p) x_release(ctx);
        ctx->peer_now = false;
        k.xCand = k.xRows0p = k.xRows0s = k.xAlive = k.xPick = 0;
        if ((rc = ensure(ctx, ctx->alive, CI)) != PFB_OK) return rc;
        if ((rc = ensure(ctx, ctx->pick, CI)) != PFB_OK) return rc;
    }
    k.xEpoch = ctx->epoch;
    if ((rc = ensure(ctx, ctx->idTm, CI * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->idInfo, CI * 2)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->pos32, CI * nG * 4)) != PFB_OK) return rc;
    if (nG > 1) {
        if ((rc = ensure(ctx, ctx->at, NA * 8)) != PFB_OK) return rc;
        if ((rc = ensure(ctx, ctx->ovf, ovfCap * 64)) != PFB_OK) return rc;
    }
    if (ctx->sharded_now && !ctx->peer_now) {
        if ((rc = ensure(ctx, ctx->candList, (size_t)nR * (ctx->capCand + 4) * 4)) != PFB_OK) return rc;
        if ((rc = ensure(ctx, ctx->rows0p, CS * 4)) != PFB_OK) return rc;
    }
    if (ctx->stream_out && (rc = host_buffers(ctx)) != PFB_OK) return rc;
    k.seq2 = (uint64_t *)ctx->seq2.p; k.nmask = (uint32_t *)ctx->nmask.p; k.wordContig = (uint32_t *)ctx->wordContig.p;
    k.sk = (uint32_t *)ctx->sk.p; k.br = (uint2 *)ctx->br.p;
    k.filter = (uint32_t *)ctx->filter.p; k.fterm = (const double *)ctx->fterm.p;
    k.tmb = (const float4 *)((const double *)ctx->fterm.p + TMB_OFF);
    k.cnt = (uint32_t *)ctx->misc.p;
    k.emitMask = (uint32_t *)ctx->emitMask.p; k.emitOff = (uint32_t *)ctx->emitOff.p;
    k.rows0 = (unsigned long long *)ctx->rows0.p; k.remap = (uint32_t *)ctx->remap.p; k.rows0p = (uint32_t *)ctx->rows0p.p;
    k.rows = (unsigned long long *)ctx->rows.p; k.key1 = (uint64_t *)ctx->key1.p; k.idMeta = (uint32_t *)ctx->idMeta.p;
    k.alive = (uint8_t *)ctx->alive.p; k.pick = (uint8_t *)ctx->pick.p;
    DevBuf aliveBuf = ctx->alive, pickBuf = ctx->pick;
    if (ctx->peer_now) {
        uint8_t *xa = (uint8_t *)ctx->xa.p;
        k.peer[ctx->rank] = xa;
        k.alive = xa + k.xAlive; k.pick = xa + k.xPick; k.rows0p = (uint32_t *)(xa + k.xRows0p);
        aliveBuf.p = k.alive; aliveBuf.cap = al256(CI); pickBuf.p = k.pick; pickBuf.cap = al256(CI);
    }
    k.idTm = (double *)ctx->idTm.p; k.idInfo = (uint16_t *)ctx->idInfo.p; k.pos32 = (uint32_t *)ctx->pos32.p;
    k.at = (uint2 *)ctx->at.p; k.ovf = (uint32_t *)ctx->ovf.p; k.ovfCount = k.cnt + C_OVF; k.ovfCap = (uint32_t)ovfCap;
    k.atWide = ctx->at_wide;
    k.capSlots = (uint32_t)CS; k.capIds = (uint32_t)CI;
    k.candList = (uint32_t *)ctx->candList.p; k.capCand = ctx->capCand; k.nRanks = (uint32_t)nR; k.rank = (uint32_t)ctx->rank;
    k.first_pass = 0;

    CK(cudaEventRecord(ctx->ev[2], st));
    if (!ctx->fterm_uploaded) {   // the parameters of a context never change: once (a small H2D copy per run would
                                  // queue up behind the 5 MB chunks of a pipelined run)
        CK(cudaMemcpyAsync(ctx->fterm.p, fterm.data(), FTERM_DOUBLES * 8, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        ctx->fterm_uploaded = true;
    }
    ZeroBatch zb;
    if ((rc = zb.add(ctx, ctx->misc, 4096)) != PFB_OK) return rc;
    if ((rc = zb.add(ctx, ctx->sk, (size_t)R.NB * 2 * 4)) != PFB_OK) return rc;
    // (peer mode: nothing of the arena is touched before every rank is done with the previous run's -- its final merge
    // reads the others' pick bytes; in a steady loop that has long happened)
    if (ctx->peer_now && ctx->epoch > 1) { k_xwait<<<1, 32, 0, st>>>(k, XP_DONE, ctx->epoch - 1); ctx->launches++; }
    if ((rc = zb.add(ctx, aliveBuf, CI)) != PFB_OK) return rc;
    if ((rc = zb.add(ctx, pickBuf, CI)) != PFB_OK) return rc;
    if ((rc = zb.launch(ctx, st)) != PFB_OK) return rc;
    if (ctx->sharded_now) {      // every segment starts with its count
        if (ctx->peer_now) CK(cudaMemsetAsync((uint8_t *)ctx->xa.p + k.xCand, 0, 16, st));
        else
            for (int r = 0; r < nR; r++)
                CK(cudaMemsetAsync((uint32_t *)ctx->candList.p + (size_t)r * (ctx->capCand + 4), 0, 16, st));
    }
    if (piped) CK(cudaStreamWaitEvent(st, ctx->evCopy[0], 0));      // only genome 1 has to be there before the key table can be built
    if (R.W1) { k_encode<<<nblk(R.W1, 256), 256, 0, st>>>(k, 0, R.W1); ctx->launches++; }
    if (!R.partial && (rc = pack_aside(ctx)) != PFB_OK) return rc;
    CK(cudaEventRecord(ctx->ev[3], st));
    trace_mark(ctx, st, "genome 1 encoded");
    // genome 1: candidates -> key table
    if (R.W1) {
        const uint32_t nCta = nblk(R.N1, CAND_POS);
        if (ctx->sharded_now) {
            const uint32_t per = (nCta + nR - 1) / nR;
            for (int r = 0; r < nR; r++) {        // this context's rank -- or, emulating, every rank in turn
                if (ctx->emulate_ranks <= 1 && r != ctx->rank) continue;
                const uint32_t c0 = std::min(nCta, per * (uint32_t)r), c1 = std::min(nCta, c0 + per);
                k.rank = (uint32_t)r;
                if (c1 > c0) { k_cand1<true><<<c1 - c0, CAND_POS + 32, 0, st>>>(k, R.N1, c0); ctx->launches++; }
            }
            k.rank = (uint32_t)ctx->rank;
        } else {
            k_cand1<false><<<nCta, CAND_POS + 32, 0, st>>>(k, R.N1, 0); ctx->launches++;
        }
    }
    CK(cudaGetLastError());
    return PFB_OK;
}

// peer mode: producer side of an exchange (the data is final on this stream) and consumer side (every rank's is)
static void x_signal(pfb_ctx *ctx, int phase) {      // (the signal itself goes out with the wait: k_xsync)
    cudaEventRecord(ctx->evX[phase < XP_ROWS_B ? phase : phase - 1][0], ctx->stream);
    ctx->xUsed[phase < XP_ROWS_B ? phase : phase - 1] = true;
}
static void x_wait(pfb_ctx *ctx, int phase) {
    trace_mark(ctx, ctx->stream, "exchange " + std::to_string(phase) + ": data ready, signalling and waiting for the peers");
    k_xsync<<<1, 32, 0, ctx->stream>>>(ctx->k, phase); ctx->launches++;
    trace_mark(ctx, ctx->stream, "exchange " + std::to_string(phase) + ": every peer has signalled");
}
static void x_done(pfb_ctx *ctx, int phase) {
    cudaEventRecord(ctx->evX[phase < XP_ROWS_B ? phase : phase - 1][1], ctx->stream);
    trace_mark(ctx, ctx->stream, "exchange " + std::to_string(phase) + ": merged");
}

// the ranks map each other's arenas (after any of them has been (re)allocated -- all of them are then, the
// capacities being the same everywhere).  One process: peer access + the pointers; one process per GPU: the IPC
// handles travel through an NCCL all-gather.  If the arenas cannot be mapped the ranks agree to use NCCL
// collectives instead and the attempt is repeated.
static int x_connect(std::vector<pfb_ctx *> &C, bool &fallback) {
    fallback = false;
    bool any = false;
    for (pfb_ctx *c : C) any |= c->peer_now && c->x_dirty;
    if (!any) return PFB_OK;
    bool ok = true;
    if (C.size() > 1) {
        for (pfb_ctx *c : C) {
            cudaSetDevice(c->device);
            for (pfb_ctx *d : C) {
                if (d == c) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, c->device, d->device) != cudaSuccess || !can) { ok = false; continue; }
                cudaError_t e = cudaDeviceEnablePeerAccess(d->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
        }
        if (ok)
            for (pfb_ctx *c : C)
                for (pfb_ctx *d : C) { c->xPeer[d->rank] = (uint8_t *)d->xa.p; c->k.peer[d->rank] = (uint8_t *)d->xa.p; }
    } else {
        pfb_ctx *ctx = C[0];
        std::string why;
        nccl::Api *a = nccl_api(why);
        if (!a) return fail(ctx, PFB_ECUDA, why);
        CK(cudaSetDevice(ctx->device));
        const int n = ctx->nRanks;
        int rc;
        if ((rc = ensure(ctx, ctx->xh, 64 * 16 + 64)) != PFB_OK) return rc;
        cudaIpcMemHandle_t mine;
        std::vector<cudaIpcMemHandle_t> all(n);
        bool mine_ok = cudaIpcGetMemHandle(&mine, ctx->xa.p) == cudaSuccess;
        if (!mine_ok) { memset(&mine, 0, sizeof mine); cudaGetLastError(); }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        CK(cudaMemcpyAsync((uint8_t *)ctx->xh.p + 64 * ctx->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
        int nrc = a->AllGather((uint8_t *)ctx->xh.p + 64 * ctx->rank, ctx->xh.p, 64, nccl::U8, ctx->comm, ctx->stream);
        if (nrc) return fail(ctx, PFB_ECUDA, std::string("NCCL: ") + a->GetErrorString(nrc));
        CK(cudaMemcpyAsync(all.data(), ctx->xh.p, 64 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ok = mine_ok;
        for (int r = 0; r < n && ok; r++) {
            if (r == ctx->rank) { ctx->xPeer[r] = (uint8_t *)ctx->xa.p; continue; }
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
            ctx->xPeer[r] = (uint8_t *)ptr; ctx->xOpened[r] = true;
        }
        // every rank must have mapped every arena, or none uses them
        uint8_t flag = ok ? 1 : 0;
        CK(cudaMemcpyAsync(ctx->xh.p, &flag, 1, cudaMemcpyHostToDevice, ctx->stream));
        nrc = a->AllReduce(ctx->xh.p, ctx->xh.p, 1, nccl::U8, nccl::MIN, ctx->comm, ctx->stream);
        if (nrc) return fail(ctx, PFB_ECUDA, std::string("NCCL: ") + a->GetErrorString(nrc));
        CK(cudaMemcpyAsync(&flag, ctx->xh.p, 1, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ok = flag != 0;
        if (ok) for (int r = 0; r < n; r++) ctx->k.peer[r] = ctx->xPeer[r];
    }
    for (pfb_ctx *c : C) c->x_dirty = false;
    if (!ok) {
        for (pfb_ctx *c : C) {
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            c->epoch = 0;                  // (nobody has signalled anything into anybody's arena)
            x_release(c);
            c->xchg_pref = 0;
        }
        fallback = true;
    }
    return PFB_OK;
}

// the filtered scan of the words [wa, wb): the second formulation for the lengths it was written for (k <= 20 and a table
// prime below 2^24: the 32-bit modular reduction), the first one for the longer lengths, where it is the faster of the
// two (config 5, k 12-30: 14.6 ms against 17.7 ms with the second formulation for every length)
static void launch_filtered_scan(pfb_ctx *ctx, bool first, uint32_t wa, uint32_t wb) {
    K &k = ctx->k;
    cudaStream_t st = ctx->stream;
    int split = 0;
    if (ctx->scan_kernel != 1 && k.P < (1u << 24)) split = std::max(0, std::min(k.nk, 21 - k.kmin));
    k.kiSplit = split;
    if (split > 0) {
        if (first) k_scan_filtered2<true><<<ctx->num_sms, 1024, SCAN2_SMEM, st>>>(k, wa, wb);
        else k_scan_filtered2<false><<<ctx->num_sms, 1024, SCAN2_SMEM, st>>>(k, wa, wb);
        ctx->launches++;
    }
    if (split < k.nk) { k_scan_filtered<<<ctx->num_sms, 1024, SCANF_SMEM, st>>>(k, wa, wb); ctx->launches++; }
}

static int phaseB(pfb_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    pfb_ctx::Run &R = ctx->run;
    const pfb_params &P = ctx->prm;
    const int nk = k.nk;
    int rc;
    if (ctx->sharded_now && ctx->peer_now) {
        x_signal(ctx, XP_CANDS);      // (here, not at the end of phase A: the peers' arenas are mapped between the two)
        x_wait(ctx, XP_CANDS);
        k_apply_cands<true><<<dim3(nblk(ctx->capCand, 256), k.nRanks), 256, 0, st>>>(k); ctx->launches++;
        x_done(ctx, XP_CANDS);
    } else if (ctx->sharded_now) { k_apply_cands<false><<<dim3(nblk(ctx->capCand, 256), k.nRanks), 256, 0, st>>>(k); ctx->launches++; }
    k_candbits<<<nblk(R.NB, 256), 256, 0, st>>>(k); ctx->launches++;
    if ((rc = device_exscan<SV_POPC, 2>(ctx, k.br, (uint32_t *)k.br + 1, R.NB, k.cnt + C_NSLOTS, nullptr)) != PFB_OK) return rc;
    ZeroBatch zb;
    if ((rc = zb.add(ctx, ctx->rows0, (size_t)ctx->capSlots * 8)) != PFB_OK) return rc;
    if (R.W1 && (rc = zb.add(ctx, ctx->emitMask, (size_t)R.W1 * 32 * 4 + 4)) != PFB_OK) return rc;   // filled by k_alive1 after genome 1's pass
    // the filter pays off while it stays sparse (expected fill 1 - exp(-keys_per_k / FILTER_BITS)) and the launch
    // is long enough to amortise the nk filter loads.  Genome 1's own pass meets ALL its candidate keys (a dense
    // filter) in one short launch: measured, the plain L2 probe is faster there at every genome size tried
    // (5 Mbp: 0.14 vs 0.18 ms), so the filter is only used for it when forced (scan mode 2, tests)
    R.use_filter1 = ctx->scan_mode == 2;
    ctx->used_filter = R.use_filter1;
    if (R.use_filter1 && (rc = zb.add(ctx, ctx->filter, (size_t)nk * FILTER_BYTES)) != PFB_OK) return rc;
    if ((rc = zb.launch(ctx, st)) != PFB_OK) return rc;
    if (R.use_filter1) { k_build_filter<<<nblk(R.NB, 256), 256, 0, st>>>(k); ctx->launches++; }
    CK(cudaEventRecord(ctx->ev[4], st));
    trace_mark(ctx, st, "key table built");
    if (R.W1) {
        // genome 1 first: its rows are indexed by slot.  The keys that collide inside genome 1 itself
        // (~1 - exp(-L / P) of them) are dead already; the survivors get an id in the reference's dict order,
        // the rows of every other genome are indexed by id, and a filter rebuilt without the dead keys sends
        // fewer windows of the other genomes to L2
        k.first_pass = ctx->sharded_now ? 2 : 1;
        const int nR = ctx->sharded_now ? ctx->eff_ranks() : 1;
        for (int r = 0; r < nR; r++) {            // this context's share -- or, emulating, every share in turn
            if (ctx->sharded_now && ctx->emulate_ranks <= 1 && r != ctx->rank) continue;
            const uint32_t wa = (uint32_t)((uint64_t)R.W1 * r / nR), wb = (uint32_t)((uint64_t)R.W1 * (r + 1) / nR);
            if (wb > wa) {
                if (R.use_filter1) { launch_filtered_scan(ctx, true, wa, wb); ctx->launches--; }
                else k_scan_plain<false><<<nblk(wb - wa, 256), 256, 0, st>>>(k, wa, wb);
                ctx->launches++;
            }
        }
        if (ctx->sharded_now) { k_rows0_pack<<<nblk(ctx->capSlots, 256), 256, 0, st>>>(k); ctx->launches++; }
        if (ctx->sharded_now && ctx->peer_now) x_signal(ctx, XP_ROWS_A);
    }
    (void)P;
    CK(cudaGetLastError());
    return PFB_OK;
}

// one batch of the other genomes: [ga, gb) are packed (unless they are already), scanned against the key table,
// judged, and their places leave for the host
static int scan_batch(pfb_ctx *ctx, size_t ga, size_t gb, bool encode, bool use_filter) {
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    pfb_ctx::Run &R = ctx->run;
    const pfb_params &P = ctx->prm;
    const size_t nG = ctx->genomes.size();
    const uint32_t wa = ctx->gWordStart[ga], wb = gb < nG ? ctx->gWordStart[gb] : R.W;
    const uint32_t ca = ctx->gFirstContig[ga], cb = gb < nG ? ctx->gFirstContig[gb] : (uint32_t)ctx->cLen.size();
    if (encode && wb > wa) { k_encode<<<nblk(wb - wa, 256), 256, 0, st>>>(k, wa, wb); ctx->launches++; }
    trace_mark(ctx, st, "scan starts: genomes " + std::to_string(ga) + ".." + std::to_string(gb - 1));
    if (wb > wa) {
        if (use_filter) launch_filtered_scan(ctx, false, wa, wb);
        else { k_scan_plain<false><<<nblk(wb - wa, 256), 256, 0, st>>>(k, wa, wb); ctx->launches++; }
    }
    if (cb - ca > gb - ga) {   // some genome of the range has more than one contig
        uint64_t nthr = (uint64_t)(cb - ca) * k.nk * 2 * (2 * P.k_max - 1);
        k_junction<<<nblk(nthr, 128), 128, 0, st>>>(k, ca, cb);
        ctx->launches++;
    }
    trace_mark(ctx, st, "scan done:   genomes " + std::to_string(ga) + ".." + std::to_string(gb - 1));
    return PFB_OK;
}

static int judge_batch(pfb_ctx *ctx, size_t ga, size_t gb) {
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    if (gb - ga > 65535u * ALIVE_GPT) return fail(ctx, PFB_ETOOLARGE, "more than 131070 genomes on one context");
    k_aliveN<<<dim3(nblk(ctx->capIds, 256), nblk(gb - ga, ALIVE_GPT)), 256, 0, st>>>(k, (uint32_t)ga, (uint32_t)gb); ctx->launches++;
    if (ctx->stream_out)
        return copy_out(ctx, (uint32_t *)ctx->hPos.p + ga * ctx->hostPitch, k.pos32 + ga * (size_t)ctx->capIds, (gb - ga) * (size_t)ctx->capIds * 4,
                        "places of a batch of genomes");
    return PFB_OK;
}

static int phaseC(pfb_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    K &k = ctx->k;
    pfb_ctx::Run &R = ctx->run;
    const pfb_params &P = ctx->prm;
    const int nk = k.nk;
    const size_t nG = ctx->genomes.size();
    int rc;
    if (R.W1) {
        if (ctx->sharded_now) {
            // the ranks' shares of genome 1's pass have been summed: back to the row format, then the (few)
            // strand-specific polluters of all of genome 1 on every rank
            if (ctx->peer_now) {
                // every rank sums its slice of the slots over all shares, then collects all slices
                x_wait(ctx, XP_ROWS_A);
                const uint32_t per = ((ctx->capSlots + ctx->nRanks - 1) / ctx->nRanks + 3u) & ~3u;
                k_rows0_reduce_slice<<<nblk(per / 4, 256), 256, 0, st>>>(k); ctx->launches++;
                x_wait(ctx, XP_ROWS_B);
                k_rows0_gather_unpack<<<nblk(ctx->capSlots / 4, 256), 256, 0, st>>>(k); ctx->launches++;
                x_done(ctx, XP_ROWS_A);
            } else {
                k_rows0_unpack<<<nblk(ctx->capSlots, 256), 256, 0, st>>>(k); ctx->launches++;
            }
            k.first_pass = 1;
            k_scan_plain<true><<<nblk(R.W1, 256), 256, 0, st>>>(k, 0, R.W1); ctx->launches++;
        }
        const uint32_t C1 = ctx->gFirstContig[0] + ctx->gNumContigs[0];
        if (C1 > 1) {
            uint64_t nthr = (uint64_t)C1 * nk * 2 * (2 * P.k_max - 1);
            k_junction<<<nblk(nthr, 128), 128, 0, st>>>(k, 0, C1);
            ctx->launches++;
        }
        k.first_pass = 0;
        trace_mark(ctx, st, "genome 1 scanned");
        k_alive1<<<nblk(ctx->capSlots, 256), 256, 0, st>>>(k); ctx->launches++;
        if ((rc = device_exscan<SV_POPC, 1>(ctx, k.emitMask, k.emitOff, R.N1, k.cnt + C_NIDS, nullptr)) != PFB_OK) return rc;
        trace_mark(ctx, st, "ids counted");
    }
    bool use_filter = false;
    {
        ZeroBatch zb;
        if (nG > 1) {
            // the per-genome rows: allocated by capacity, zeroed by need (the count is on the device)
            if (!R.rows_zeroed) {
                const unsigned ry = (unsigned)std::min<size_t>(nG - 1, 65535);
                k_zero_rows<<<dim3(nblk((ctx->capIds + 1) / 2, 256), ry), 256, 0, st>>>(k.rows, k.capIds, (uint32_t)(nG - 1), k.cnt + C_NIDS, k.capIds);
                ctx->launches++;
            }
            // the other genomes only meet the keys that survived genome 1: decide again whether their filter
            // stays sparse enough (long genomes lose most keys to collisions inside genome 1).  The decision must
            // not wait for the id count: it is made from the last run's count, or from an estimate
            // (expected: the typical share of candidate windows, times the share exp(-L / P) of them that are alone
            // in their bin after genome 1's own pass)
            const double ids = ctx->lastIds ? (double)ctx->lastIds
                                            : (P.prefilter ? 0.09 : 0.75) * (double)R.windows1 *
                                                  std::exp(-(double)ctx->genomes[0].n_bases / (double)P.table_size);
            use_filter = ctx->scan_mode == 2 || (ctx->scan_mode == 0 && ids / nk < 0.35 * FILTER_BITS && R.W_alloc >= 4096);
            ctx->used_filter = use_filter;
            if (use_filter && (rc = zb.add(ctx, ctx->filter, (size_t)nk * FILTER_BYTES)) != PFB_OK) return rc;
            if ((rc = zb.launch(ctx, st)) != PFB_OK) return rc;
            const size_t NA = (size_t)nk * k.PW16;
            k_build_at<<<nblk(NA, 256), 256, 0, st>>>(k, use_filter ? 1 : 0); ctx->launches++;
            trace_mark(ctx, st, "key table of the other genomes built");
        } else if (R.W1) {
            k_assign_ids<<<nblk(ctx->capSlots, 256), 256, 0, st>>>(k); ctx->launches++;
        }
    }
    // what is known per id once the ids exist: length, GC, Tm, filter flags (and genome 1's places)
    k_id_records<<<nblk(ctx->capIds, 256), 256, 0, st>>>(k); ctx->launches++;
    if (ctx->stream_out) {
        if (ctx->fetch_records) {
            if ((rc = copy_out(ctx, ctx->hWord.p, k.key1, (size_t)ctx->capIds * 8, "words")) != PFB_OK) return rc;
            CK(cudaMemcpyAsync(ctx->hTm.p, k.idTm, (size_t)ctx->capIds * 8, cudaMemcpyDeviceToHost, ctx->d2hStream));
            CK(cudaMemcpyAsync(ctx->hInfo.p, k.idInfo, (size_t)ctx->capIds * 2, cudaMemcpyDeviceToHost, ctx->d2hStream));
            CK(cudaMemcpyAsync(ctx->hPos.p, k.pos32, (size_t)ctx->capIds * 4, cudaMemcpyDeviceToHost, ctx->d2hStream));
            if (ctx->tracing) trace_mark(ctx, ctx->d2hStream, "result copy landed: Tm, info, places in genome 1");
        } else if ((rc = copy_out(ctx, ctx->hPos.p, k.pos32, (size_t)ctx->capIds * 4)) != PFB_OK) return rc;
    }
    if (R.partial) {
        // the other FASTA files have been parsed behind their copies: complete the layout
        if ((rc = fasta_finish(ctx)) != PFB_OK) return rc;
        R.W = ctx->totalWords;
        if ((rc = ensure_keep(ctx, ctx->seq2, (size_t)(R.W + 2) * 8, (size_t)(R.W1 + 2) * 8)) != PFB_OK) return rc;
        if ((rc = ensure_keep(ctx, ctx->nmask, (size_t)(R.W + 1) * 4, (size_t)(R.W1 + 1) * 4)) != PFB_OK) return rc;
        if ((rc = ensure_keep(ctx, ctx->wordContig, (size_t)(R.W + 1) * 4, (size_t)(R.W1 + 1) * 4)) != PFB_OK) return rc;
        k.seq2 = (uint64_t *)ctx->seq2.p; k.nmask = (uint32_t *)ctx->nmask.p; k.wordContig = (uint32_t *)ctx->wordContig.p;
    }
    if (nG > 1) {
        if ((rc = pack_aside(ctx)) != PFB_OK) return rc;          // (FASTA files parsed late: only now)
        if (!R.piped) {
            CK(cudaStreamWaitEvent(st, ctx->evEnc[0], 0));
            CK(cudaEventRecord(ctx->ev[7], st));
            if ((rc = scan_batch(ctx, 1, nG, false, use_filter)) != PFB_OK) return rc;
            CK(cudaEventRecord(ctx->ev[8], st));
            CK(cudaEventRecord(ctx->ev[5], st));
            if ((rc = judge_batch(ctx, 1, nG)) != PFB_OK) return rc;
        } else {
            // the other genomes are packed and scanned in batches behind their H2D copies (which run on their own
            // stream, see upload_impl): the stream waits for a batch's last copy, the host does not.  Where to cut
            // is planned here, before anything runs, from rough rates (a launch over few genomes is inefficient --
            // nk filter loads, one word per thread -- so batches must not be small; a batch cannot start before its
            // last copy has landed; the places of a batch leave for the host while the next one is scanned)
            size_t g = 1;
            for (size_t b = 0; b < R.cuts.size(); b++) {
                const size_t g2 = R.cuts[b];
                CK(cudaStreamWaitEvent(st, ctx->evEnc[b], 0));
                if ((rc = scan_batch(ctx, g, g2, false, use_filter)) != PFB_OK) return rc;
                if ((rc = judge_batch(ctx, g, g2)) != PFB_OK) return rc;
                g = g2;
            }
            CK(cudaEventRecord(ctx->ev[5], st));
        }
    } else {
        if (R.piped) CK(cudaStreamWaitEvent(st, ctx->evCopy[nG - 1], 0));
        CK(cudaEventRecord(ctx->ev[5], st));
    }
    if (ctx->peer_now) x_signal(ctx, XP_ALIVE);
    trace_mark(ctx, st, "all genomes scanned and judged");
    CK(cudaGetLastError());
    return PFB_OK;
}

static int phaseD(pfb_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->peer_now) {      // alive = MIN over the ranks, read from their arenas
        x_wait(ctx, XP_ALIVE);
        k_bytes_merge<false><<<nblk(ctx->capIds / 4, 256), 256, 0, ctx->stream>>>(ctx->k, ctx->k.xAlive); ctx->launches++;
        x_done(ctx, XP_ALIVE);
    }
    k_pick<<<nblk(ctx->capIds, 256), 256, 0, ctx->stream>>>(ctx->k); ctx->launches++;
    if (ctx->peer_now) x_signal(ctx, XP_PICK);
    CK(cudaGetLastError());
    return PFB_OK;
}

static int phaseE(pfb_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if (ctx->peer_now) {      // pick = MAX over the ranks, the counts, "done with the peers' arenas": one launch
        x_wait(ctx, XP_PICK);
        k_finish<true><<<ctx->num_sms * 2, 256, 0, st>>>(ctx->k, ctx->countsDev); ctx->launches++;
        x_done(ctx, XP_PICK);
    } else {
        k_finish<false><<<ctx->num_sms * 2, 256, 0, st>>>(ctx->k, ctx->countsDev); ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[6], st));
    if (ctx->stream_out) {
        if ((rc = copy_out(ctx, ctx->hAlive.p, ctx->k.alive, ctx->capIds, "alive")) != PFB_OK) return rc;
        CK(cudaMemcpyAsync(ctx->hPick.p, ctx->k.pick, ctx->capIds, cudaMemcpyDeviceToHost, ctx->d2hStream));
        if (ctx->tracing) trace_mark(ctx, ctx->d2hStream, "result copy landed: pick (the last byte)");
    }
    trace_mark(ctx, st, "results final (end of the stage)");
    CK(cudaGetLastError());
    return PFB_OK;
}

// the collectives between the phases; C = the contexts this host thread drives
enum { COLL_CANDS = 0, COLL_ROWS0 = 1, COLL_ALIVE = 2, COLL_PICK = 3 };

static int collective(std::vector<pfb_ctx *> &C, int which) {
    bool any = false;
    for (pfb_ctx *c : C) any |= c->comm && c->nRanks > 1 && !c->peer_now;
    if (!any) return PFB_OK;
    std::string why;
    nccl::Api *a = nccl_api(why);
    if (!a) return fail(C[0], PFB_ECUDA, why);
    for (pfb_ctx *c : C) {
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, PFB_ECUDA, "cudaSetDevice failed");
        cudaEventRecord(c->evX[which][0], c->stream);      // the two events bracket this collective on the context's stream
        c->xUsed[which] = true;
    }
    int rc = 0;
    if (C.size() > 1) a->GroupStart();
    for (pfb_ctx *c : C) {
        if (!(c->comm && c->nRanks > 1) || c->peer_now) continue;
        cudaSetDevice(c->device);
        K &k = c->k;
        switch (which) {
        case COLL_CANDS:
            if (c->sharded_now)
                rc = a->AllGather(k.candList + (size_t)c->rank * (c->capCand + 4), k.candList, c->capCand + 4, nccl::U32, c->comm, c->stream);
            break;
        case COLL_ROWS0:
            if (c->sharded_now && c->run.W1) rc = a->AllReduce(k.rows0p, k.rows0p, c->capSlots, nccl::U32, nccl::SUM, c->comm, c->stream);
            break;
        case COLL_ALIVE:
            rc = a->AllReduce(k.alive, k.alive, c->capIds, nccl::U8, nccl::MIN, c->comm, c->stream);
            break;
        case COLL_PICK:
            rc = a->AllReduce(k.pick, k.pick, c->capIds, nccl::U8, nccl::MAX, c->comm, c->stream);
            break;
        }
        if (rc) break;
    }
    if (C.size() > 1) { int rc2 = a->GroupEnd(); if (!rc) rc = rc2; }
    if (rc) return fail(C[0], PFB_ECUDA, std::string("NCCL: ") + a->GetErrorString(rc));
    for (pfb_ctx *c : C) {
        cudaSetDevice(c->device);
        cudaEventRecord(c->evX[which][1], c->stream);
    }
    return PFB_OK;
}

// end of a run on one context: wait, read the counters, decide whether the capacities were enough
static int finish(pfb_ctx *ctx, bool &again) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    const uint32_t nslots = ctx->countsHost[C_NSLOTS], nids = ctx->countsHost[C_NIDS], flags = ctx->countsHost[C_FLAGS];
    ctx->nslots = nslots; ctx->nids = nids;
    ctx->nrec = ctx->countsHost[C_NALIVE]; ctx->npick = ctx->countsHost[C_NPICK];
    auto round256 = [](uint64_t x) { return (uint32_t)std::min<uint64_t>((x + 255) & ~255ull, 0xFFFFFF00ull); };
    if (flags & OVF_PEER_TIMEOUT) return fail(ctx, PFB_ECUDA, "a rank did not reach an exchange within the time limit (peer arenas)");
    if (flags) {
        // a count did not fit: grow what overflowed and repeat (the slot count is exact whatever happened; the id
        // count only when every slot had its row)
        if ((flags & OVF_SLOTS) || nslots > ctx->capSlots) {
            ctx->capSlots = round256((uint64_t)nslots + nslots / 16 + 1024);
            ctx->capIds = std::max(ctx->capIds, ctx->capSlots);
        } else if (flags & OVF_IDS) {
            ctx->capIds = round256((uint64_t)nids + nids / 16 + 1024);
        }
        if (flags & OVF_LISTS) ctx->at_wide = std::min(ctx->at_wide, 0xFFFFFFu);   // (the list capacity follows capIds)
        if (flags & OVF_CAND) {            // a rank's candidate list was longer than planned (every rank sees every list's length)
            const uint32_t need = ctx->countsHost[C_MAXCAND];
            ctx->capCandMin = round256((uint64_t)need + need / 16 + 1024);
        }
        ctx->lastSlots = ctx->lastIds = 0;
        ctx->reruns++;
        again = true;
        return PFB_OK;
    }
    ctx->lastSlots = nslots; ctx->lastIds = nids;
    pfb_timing &t = ctx->tm;
    cudaEventElapsedTime(&t.ms_encode, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&t.ms_first, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&t.ms_scan, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&t.ms_finalize, ctx->ev[5], ctx->ev[6]);
    t.ms_scan_main = 0.f;
    t.scan_main_bases = 0;
    if (!ctx->run.piped && ctx->genomes.size() > 1 && cudaEventElapsedTime(&t.ms_scan_main, ctx->ev[7], ctx->ev[8]) == cudaSuccess) {
        for (size_t g = 1; g < ctx->genomes.size(); g++) t.scan_main_bases += ctx->genomes[g].n_bases;
    } else {
        t.ms_scan_main = 0.f;
        cudaGetLastError();
    }
    cudaEventElapsedTime(&t.ms_total, ctx->ev[2], ctx->ev[6]);
    t.ms_exchange = 0.f;
    for (int x = 0; x < 4; x++) {
        float ms = 0.f;
        if (ctx->xUsed[x] && cudaEventElapsedTime(&ms, ctx->evX[x][0], ctx->evX[x][1]) == cudaSuccess) t.ms_exchange += ms;
        ctx->xUsed[x] = false;
    }
    cudaGetLastError();
    uint64_t nb = 0, nw = 0;
    for (auto &g : ctx->genomes) {
        nb += g.n_bases;
        for (size_t c = 0; c + 1 < g.off.size(); c++) {
            uint64_t len = g.off[c + 1] - g.off[c];
            for (int kk = ctx->prm.k_min; kk <= ctx->prm.k_max; kk++) if (len >= (uint64_t)kk) nw += len - kk + 1;
        }
    }
    t.n_bases = nb; t.n_windows = nw; t.n_slots = ctx->nslots; t.n_records = ctx->nrec; t.n_picked = ctx->npick;
    t.n_ids = ctx->nids;
    t.n_launches = ctx->launches + ctx->fa_launches;
    t.used_filter = ctx->used_filter ? 1 : 0;
    t.sharded_first_genome = ctx->sharded_now ? 1 : 0;
    t.n_reruns = (uint64_t)ctx->reruns;
    t.exchange_mode = ctx->nRanks > 1 && ctx->comm ? (ctx->peer_now ? 2 : 1) : 0;
    ctx->computed = true;
    ctx->host_valid = ctx->stream_out;
    return PFB_OK;
}

// all contexts driven by this host thread through one run (see the phase list above)
static int run_group(std::vector<pfb_ctx *> &C, bool piped) {
    int rc;
    for (int attempt = 0; attempt < 6; attempt++) {
        for (pfb_ctx *c : C) if ((rc = phaseA(c, piped)) != PFB_OK) return rc;
        bool fallback = false;
        if ((rc = x_connect(C, fallback)) != PFB_OK) return rc;
        if (fallback) continue;       // the arenas could not be mapped: the same run again with NCCL collectives
        if ((rc = collective(C, COLL_CANDS)) != PFB_OK) return rc;
        for (pfb_ctx *c : C) if ((rc = phaseB(c)) != PFB_OK) return rc;
        if ((rc = collective(C, COLL_ROWS0)) != PFB_OK) return rc;
        for (pfb_ctx *c : C) if ((rc = phaseC(c)) != PFB_OK) return rc;
        if ((rc = collective(C, COLL_ALIVE)) != PFB_OK) return rc;
        for (pfb_ctx *c : C) if ((rc = phaseD(c)) != PFB_OK) return rc;
        if ((rc = collective(C, COLL_PICK)) != PFB_OK) return rc;
        for (pfb_ctx *c : C) if ((rc = phaseE(c)) != PFB_OK) return rc;
        bool again = false;
        for (pfb_ctx *c : C) if ((rc = finish(c, again)) != PFB_OK) return rc;
        if (!again) return PFB_OK;
        // every rank of a job sees the same counts (genome 1 is everywhere, the candidate lists are gathered), so
        // all of them come here together.  The inputs are resident by now: the repeat does not wait for copies
        if (piped) {
            for (pfb_ctx *c : C) {
                cudaSetDevice(c->device);
                cudaStreamSynchronize(c->copyStream);
                cudaStreamSynchronize(c->parseStream);
                if (c->laid_out < c->genomes.size()) return fail(c, PFB_ESTATE, "layout incomplete after a pipelined run");
            }
            piped = false;
        }
        for (pfb_ctx *c : C) { cudaSetDevice(c->device); cudaStreamSynchronize(c->d2hStream); }
    }
    return fail(C[0], PFB_ECUDA, "the buffer capacities did not settle");
}

int pfb_compute(pfb_ctx *ctx) {
    if (!ctx) return PFB_EINVAL;
    if (!ctx->uploaded) return fail(ctx, PFB_ESTATE, "pfb_upload has not run");
    ctx->stream_out = false;
    ctx->host_valid = false;
    std::vector<pfb_ctx *> C{ctx};
    int rc = run_group(C, ctx->pipelined);
    if (ctx->tracing) trace_dump(ctx);
    return rc;
}

int pfb_run(pfb_ctx *ctx) {
    // H2D of genome i+1.. overlaps the key-table build and the scan of the genomes already resident, and the
    // results travel back while later genomes are still scanned
    int rc = upload_checked(ctx, true);
    if (rc != PFB_OK) return rc;
    ctx->stream_out = true;
    ctx->host_valid = false;
    std::vector<pfb_ctx *> C{ctx};
    rc = run_group(C, true);
    int rc2 = pfb_upload_wait(ctx);             // the host buffers are borrowed until here, success or not
    if (rc == PFB_OK && rc2 == PFB_OK) { if (cudaStreamSynchronize(ctx->d2hStream) != cudaSuccess) rc = fail(ctx, PFB_ECUDA, "result copies failed"); }
    if (ctx->tracing) trace_dump(ctx);
    return rc != PFB_OK ? rc : rc2;
}

// ---- multi-GPU, one process per GPU: every rank creates its context, rank 0 makes the id, all ranks join
int pfb_comm_unique_id(uint8_t *id128) {
    if (!id128) return PFB_EINVAL;
    std::string why;
    nccl::Api *a = nccl_api(why);
    if (!a) return fail(nullptr, PFB_ECUDA, why);
    nccl::UniqueId id;
    int rc = a->GetUniqueId(&id);
    if (rc) return fail(nullptr, PFB_ECUDA, std::string("ncclGetUniqueId: ") + a->GetErrorString(rc));
    memcpy(id128, id.internal, 128);
    return PFB_OK;
}

int pfb_comm_init_rank(pfb_ctx *ctx, int n_ranks, int rank, const uint8_t *id128) {
    if (!ctx || !id128) return PFB_EINVAL;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, PFB_EINVAL, "need 0 <= rank < n_ranks");
    if (ctx->comm) return fail(ctx, PFB_ESTATE, "the context already has a communicator");
    std::string why;
    nccl::Api *a = nccl_api(why);
    if (!a) return fail(ctx, PFB_ECUDA, why);
    CK(cudaSetDevice(ctx->device));
    nccl::UniqueId id;
    memcpy(id.internal, id128, 128);
    int rc = a->CommInitRank(&ctx->comm, n_ranks, id, rank);
    if (rc) { ctx->comm = nullptr; return fail(ctx, PFB_ECUDA, std::string("ncclCommInitRank: ") + a->GetErrorString(rc)); }
    ctx->comm_owned = true;
    ctx->nRanks = n_ranks; ctx->rank = rank;
    return PFB_OK;
}

int pfb_set_option(pfb_ctx *ctx, int option, int value) {
    if (!ctx) return PFB_EINVAL;
    switch (option) {
    case PFB_OPT_SCAN_MODE: return pfb_set_scan_mode(ctx, value);
    case PFB_OPT_FETCH_RECORDS: ctx->fetch_records = value ? 1 : 0; return PFB_OK;
    case PFB_OPT_SHARD_FIRST: ctx->shard_g1 = value ? 1 : 0; return PFB_OK;
    case PFB_OPT_EMULATE_RANKS: ctx->emulate_ranks = value > 1 ? value : 0; ctx->caps_for_prefilter = -1; return PFB_OK;
    default: return fail(ctx, PFB_EINVAL, "unknown option");
    }
}

// ---- multi-GPU, one process: N contexts, one communicator per context, one host thread
int pfb_multi_create(pfb_multi **out, const int *device_ids, int n_dev, const pfb_params *params) {
    if (!out) return fail(nullptr, PFB_EINVAL, "out is NULL");
    *out = nullptr;
    if (!device_ids || n_dev < 1) return fail(nullptr, PFB_EINVAL, "need at least one device");
    for (int i = 0; i < n_dev; i++)
        for (int j = 0; j < i; j++)
            if (device_ids[i] == device_ids[j]) return fail(nullptr, PFB_EINVAL, "a device is listed twice");
    pfb_multi *m = new pfb_multi();
    for (int i = 0; i < n_dev; i++) {
        pfb_ctx *c = nullptr;
        int rc = pfb_create(&c, device_ids[i], params);
        if (rc != PFB_OK) { pfb_multi_destroy(m); return rc; }
        c->owner = m;
        m->ctx.push_back(c);
    }
    if (n_dev > 1) {
        std::string why;
        nccl::Api *a = nccl_api(why);
        if (!a) { pfb_multi_destroy(m); return fail(nullptr, PFB_ECUDA, why); }
        std::vector<nccl::Comm *> comms(n_dev, nullptr);
        int rc = a->CommInitAll(comms.data(), n_dev, device_ids);
        if (rc) { pfb_multi_destroy(m); return fail(nullptr, PFB_ECUDA, std::string("ncclCommInitAll: ") + a->GetErrorString(rc)); }
        for (int i = 0; i < n_dev; i++) {
            m->ctx[i]->comm = comms[i]; m->ctx[i]->comm_owned = true;
            m->ctx[i]->nRanks = n_dev; m->ctx[i]->rank = i;
            m->ctx[i]->fetch_records = i == 0;      // the per-id records are the same on every GPU
        }
    }
    *out = m;
    return PFB_OK;
}

void pfb_multi_destroy(pfb_multi *m) {
    if (!m) return;
    for (pfb_ctx *c : m->ctx) pfb_destroy(c);
    delete m;
}

const char *pfb_multi_last_error(const pfb_multi *m) {
    if (!m) return g_create_err.c_str();
    for (pfb_ctx *c : m->ctx) if (!c->err.empty()) return c->err.c_str();
    return m->err.c_str();
}

// genome 0 goes to every context (it defines the key table and the ids), the others round-robin
static int multi_add(pfb_multi *m, int (*add)(pfb_ctx *, const uint8_t *, uint64_t, const uint64_t *, uint32_t), const uint8_t *data,
                     uint64_t n, const uint64_t *off, uint32_t nc) {
    if (!m) return PFB_EINVAL;
    const size_t g = m->where.size();
    int rc;
    if (g == 0) {
        for (pfb_ctx *c : m->ctx) if ((rc = add(c, data, n, off, nc)) != PFB_OK) return rc;
        m->where.emplace_back(0, 0u);
    } else {
        const int ci = (int)((g - 1) % m->ctx.size());
        pfb_ctx *c = m->ctx[ci];
        if ((rc = add(c, data, n, off, nc)) != PFB_OK) return rc;
        m->where.emplace_back(ci, (uint32_t)c->genomes.size() - 1);
    }
    return PFB_OK;
}

static int add_fasta4(pfb_ctx *c, const uint8_t *d, uint64_t n, const uint64_t *, uint32_t) { return pfb_add_fasta(c, d, n); }
static int add_genbank4(pfb_ctx *c, const uint8_t *d, uint64_t n, const uint64_t *, uint32_t) { return pfb_add_genbank(c, d, n); }

int pfb_multi_add_genome(pfb_multi *m, const uint8_t *bases, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs) {
    return multi_add(m, pfb_add_genome, bases, n_bases, contig_off, n_contigs);
}
int pfb_multi_add_fasta(pfb_multi *m, const uint8_t *file_bytes, uint64_t n_bytes) { return multi_add(m, add_fasta4, file_bytes, n_bytes, nullptr, 0); }
int pfb_multi_add_genbank(pfb_multi *m, const uint8_t *file_bytes, uint64_t n_bytes) { return multi_add(m, add_genbank4, file_bytes, n_bytes, nullptr, 0); }

int pfb_multi_run(pfb_multi *m) {
    if (!m || m->ctx.empty()) return PFB_EINVAL;
    if (m->where.empty()) return fail(m->ctx[0], PFB_ESTATE, "no genome added");
    int rc = PFB_OK;
    for (pfb_ctx *c : m->ctx) {
        c->err.clear();
        if ((rc = upload_checked(c, true)) != PFB_OK) break;
        c->stream_out = true; c->host_valid = false;
    }
    if (rc == PFB_OK) rc = run_group(m->ctx, true);
    for (pfb_ctx *c : m->ctx) {
        int rc2 = pfb_upload_wait(c);
        if (rc == PFB_OK) rc = rc2;
        if (rc == PFB_OK && (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->d2hStream) != cudaSuccess))
            rc = fail(c, PFB_ECUDA, "result copies failed");
    }
    return rc;
}

int pfb_multi_size(const pfb_multi *m) { return m ? (int)m->ctx.size() : 0; }

int pfb_multi_context(pfb_multi *m, int index, pfb_ctx **out) {
    if (!m || !out || index < 0 || index >= (int)m->ctx.size()) return PFB_EINVAL;
    *out = m->ctx[index];
    return PFB_OK;
}

int pfb_multi_locate(const pfb_multi *m, uint32_t genome_idx, int *context_index, uint32_t *local_idx) {
    if (!m || genome_idx >= m->where.size()) return PFB_EINVAL;
    if (context_index) *context_index = m->where[genome_idx].first;
    if (local_idx) *local_idx = m->where[genome_idx].second;
    return PFB_OK;
}

// ---- results.  The device keeps everything indexed by id (the reference's dict order); the host gets the per-id
// columns once per run -- during the run (pfb_run, pfb_multi_run) or on the first result call (pfb_compute) -- and
// the record-oriented calls below compact them on the host.
static int fetch_host(pfb_ctx *ctx) {
    if (!ctx->computed) return fail(ctx, PFB_ESTATE, "no completed run");
    if (ctx->host_valid) return PFB_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = host_buffers(ctx)) != PFB_OK) return rc;
    const size_t CI = ctx->capIds, nG = ctx->genomes.size();
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->hWord.p, ctx->k.key1, CI * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->hTm.p, ctx->k.idTm, CI * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->hInfo.p, ctx->k.idInfo, CI * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->hPos.p, ctx->k.pos32, CI * nG * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->hAlive.p, ctx->k.alive, CI, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->hPick.p, ctx->k.pick, CI, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->host_valid = true;
    return PFB_OK;
}

// records stayed on the device (fetch_records = 0) and are asked for after all: get them now
static int fetch_records_now(pfb_ctx *ctx) {
    if (ctx->fetch_records || !ctx->stream_out) return PFB_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t CI = ctx->capIds;
    CK(cudaMemcpyAsync(ctx->hWord.p, ctx->k.key1, CI * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->hTm.p, ctx->k.idTm, CI * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->hInfo.p, ctx->k.idInfo, CI * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PFB_OK;
}

int pfb_result_counts(pfb_ctx *ctx, uint64_t *n_records, uint64_t *n_picked) {
    if (!ctx) return PFB_EINVAL;
    if (!ctx->computed) return fail(ctx, PFB_ESTATE, "no completed run");
    if (n_records) *n_records = ctx->nrec;
    if (n_picked) *n_picked = ctx->npick;
    return ctx->nrec ? PFB_OK : PFB_ENOSHARED;
}

int pfb_result_by_id(pfb_ctx *ctx, pfb_idview *out) {
    if (!ctx || !out) return PFB_EINVAL;
    int rc = fetch_host(ctx);
    if (rc != PFB_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->d2hStream));
    out->n_ids = ctx->nids; out->pitch = ctx->hostPitch; out->n_genomes = ctx->genomes.size();
    out->records_valid = (ctx->fetch_records || !ctx->stream_out) ? 1 : 0;
    out->word = (const uint64_t *)ctx->hWord.p; out->tm = (const double *)ctx->hTm.p; out->info = (const uint16_t *)ctx->hInfo.p;
    out->pos = (const uint32_t *)ctx->hPos.p; out->alive = (const uint8_t *)ctx->hAlive.p; out->pick = (const uint8_t *)ctx->hPick.p;
    return PFB_OK;
}

static inline pfb_record make_record(const pfb_ctx *ctx, uint64_t id) {
    const uint16_t info = ((const uint16_t *)ctx->hInfo.p)[id];
    pfb_record r;
    r.word = ((const uint64_t *)ctx->hWord.p)[id];
    r.tm = ((const double *)ctx->hTm.p)[id];
    r.slot = (uint32_t)id;
    r.gc_count = (uint16_t)((info >> 5) & 63u);
    r.k = (uint8_t)((info & 31u) + 1u);
    r.flags = (uint8_t)((info >> 11) | (((const uint8_t *)ctx->hPick.p)[id] ? PFB_F_PICK : 0u));
    return r;
}

// genome-local padded coordinate -> (contig index within the genome, start within the contig)
static inline void decode_place(const pfb_ctx *ctx, uint32_t g, uint32_t lp, uint32_t &contig, uint32_t &start) {
    const uint32_t c0 = ctx->gFirstContig[g], nc = ctx->gNumContigs[g];
    if (nc == 1) { contig = 0; start = lp; return; }
    const uint32_t w = ctx->gWordStart[g] + (lp >> 5);
    uint32_t lo = c0, hi = c0 + nc;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ctx->cWordStart[mid] <= w) lo = mid; else hi = mid;
    }
    while (lo + 1 < c0 + nc && ctx->cWordStart[lo + 1] <= w) lo++;          // (empty contigs share their first word)
    contig = lo - c0;
    start = lp - 32u * (ctx->cWordStart[lo] - ctx->gWordStart[g]);
}

int pfb_result_records(pfb_ctx *ctx, pfb_record *out, uint64_t cap) {
    if (!ctx || (!out && cap)) return PFB_EINVAL;
    int rc = fetch_host(ctx);
    if (rc != PFB_OK) return rc;
    if ((rc = fetch_records_now(ctx)) != PFB_OK) return rc;
    if (cap < ctx->nrec) return fail(ctx, PFB_EINVAL, "record buffer too small");
    const uint8_t *alive = (const uint8_t *)ctx->hAlive.p;
    uint64_t n = 0;
    for (uint64_t id = 0; id < ctx->nids; id++)
        if (alive[id]) out[n++] = make_record(ctx, id);
    return PFB_OK;
}

int pfb_result_positions(pfb_ctx *ctx, uint32_t genome_idx, int picked_only, pfb_pos *out, uint64_t cap) {
    if (!ctx || (!out && cap)) return PFB_EINVAL;
    int rc = fetch_host(ctx);
    if (rc != PFB_OK) return rc;
    if (genome_idx >= ctx->genomes.size()) return fail(ctx, PFB_EINVAL, "genome index out of range");
    const uint64_t want = picked_only ? ctx->npick : ctx->nrec;
    if (cap < want) return fail(ctx, PFB_EINVAL, "position buffer too small");
    const uint8_t *sel = (const uint8_t *)(picked_only ? ctx->hPick.p : ctx->hAlive.p);
    const uint32_t *row = (const uint32_t *)ctx->hPos.p + (size_t)genome_idx * ctx->hostPitch;
    uint64_t n = 0;
    for (uint64_t id = 0; id < ctx->nids; id++) {
        if (!sel[id]) continue;
        decode_place(ctx, genome_idx, row[id] & 0x7FFFFFFFu, out[n].contig, out[n].start);
        out[n].strand = row[id] >> 31;
        n++;
    }
    return PFB_OK;
}

int pfb_result_picked(pfb_ctx *ctx, pfb_record *rec_out, uint32_t *start_strand_out, uint32_t *contig_out, uint64_t cap) {
    if (!ctx) return PFB_EINVAL;
    int rc = fetch_host(ctx);
    if (rc != PFB_OK) return rc;
    if ((rc = fetch_records_now(ctx)) != PFB_OK) return rc;
    if (cap < ctx->npick) return fail(ctx, PFB_EINVAL, "output buffers too small");
    if (!ctx->npick) return PFB_OK;
    if (!rec_out || !start_strand_out) return fail(ctx, PFB_EINVAL, "NULL output buffer");
    const size_t nG = ctx->genomes.size();
    const uint8_t *pick = (const uint8_t *)ctx->hPick.p;
    uint64_t n = 0;
    for (uint64_t id = 0; id < ctx->nids; id++) {
        if (!pick[id]) continue;
        rec_out[n] = make_record(ctx, id);
        for (size_t g = 0; g < nG; g++) {
            const uint32_t v = ((const uint32_t *)ctx->hPos.p)[g * ctx->hostPitch + id];
            uint32_t contig, start;
            decode_place(ctx, (uint32_t)g, v & 0x7FFFFFFFu, contig, start);
            start_strand_out[g * cap + n] = start | (v & 0x80000000u);
            if (contig_out) contig_out[g * cap + n] = contig;
        }
        n++;
    }
    return PFB_OK;
}

int pfb_host_alloc(void **out, uint64_t n_bytes) {
    if (!out) return PFB_EINVAL;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, n_bytes ? n_bytes : 1, cudaHostAllocDefault);
    return e == cudaSuccess ? PFB_OK : PFB_ECUDA;
}

int pfb_host_free(void *ptr) {
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) return PFB_ECUDA;
    return PFB_OK;
}

int pfb_timings(pfb_ctx *ctx, pfb_timing *out) {
    if (!ctx || !out) return PFB_EINVAL;
    *out = ctx->tm;
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
    K k = ctx->k;
    std::vector<double> fterm;
    fill_filter_constants(ctx->prm, k, fterm);
    int rc;
    if ((rc = ensure(ctx, ctx->fterm, FTERM_DOUBLES * 8)) != PFB_OK) return rc;
    if ((rc = ensure(ctx, ctx->misc, 4096)) != PFB_OK) return rc;
    k.fterm = (const double *)ctx->fterm.p;
    k.tmb = (const float4 *)((const double *)ctx->fterm.p + TMB_OFF);
    CK(cudaMemcpyAsync(ctx->fterm.p, fterm.data(), FTERM_DOUBLES * 8, cudaMemcpyHostToDevice, ctx->stream));
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
