// pfb_outgroup.cu -- B200 (sm_100a) outgroup screen of primerForge, from scratch.
//
// What the reference computes (bin/removeOutgroupPrimers.py:181-313; file:line citations are into /root/reference):
// an Aho-Corasick automaton over the distinct primer strings of all pairs (:14-31) is run over every contig of
// every outgroup genome, forward and reverse-complemented (:215-226); per contig and strand it keeps the start
// positions of every primer string (:34-62, '-' starts are len - start_in_rc - 1 = the rightmost (+) coordinate);
// per pair and contig the PCR product sizes are rStart - fStart + 1 > 0 over plus[fwd] x minus[rev] and
// plus[rev] x minus[fwd] (:65-122); a pair with a size inside params.disallowedLens is deleted (:281-287), the
// other sizes are kept as (contig, size) per outgroup genome (:289-291).
//
// How it is computed here: no automaton, no strings.  The primer strings become (k, 2-bit word) keys in one
// open-addressing table per length, entered under their CANONICAL word min(w, rc(w)) with two key indices
// (the key that reads like the canonical word, the key that reads like its reverse complement).  The outgroup
// contigs are packed to 2 bits per base (+ an invalid bit for everything that is not upper-case ACGT) and ONE pass
// rolls the forward and the reverse-complement word of every window: one probe per window and length answers
// both strands.  Hits (key, contig, start, strand) are bucketed per (key, strand) by a counting sort on the
// device and one thread per pair enumerates its products.  A table slot is 8 bytes (32-bit tag of the canonical
// word, index of its 16-byte record; load 1/3), so that a million keys still probe an L2-resident table; the record
// is only read when the tag matches.

#include "../../include/pfb200.h"
#include "pfb_devscan.cuh"
#include "pfb_bulk.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace og {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t SMEM_FILTER_BYTES = 192u * 1024u;   // the filter of every length together (k_og_scan_smem)

struct Args {   // kernel arguments (by value)
    const uint8_t *raw;            // ASCII bases of every contig, contigs concatenated
    uint64_t *seq2;                // 2 bits per base, A0 T1 C2 G3, first base most significant, contigs word-aligned
    uint32_t *nmask;               // 1 bit per base: not an upper-case nucleotide (or padding)
    const uint32_t *cWordStart, *cLen, *cGenome;
    const uint64_t *cRawStart;
    uint32_t nContigs, totalWords;
    int kmin, nk;
    const uint2 *tab;              // per length: slots {tag of the canonical word (0 = empty), record index}
    const uint4 *recs;             // {canonical word lo, hi, key reading like it, key reading like its rc}
    uint32_t tabOff[34], tabSize[33];   // slot range of length kmin + ki
    const uint32_t *bits;          // per length: the filter, bit umulhi(hash, nBits) set for every record
    uint32_t bitOff[33], nBits[33];     // (word offset, bits) of length kmin + ki
    uint32_t bitWords;
    uint4 *hits;                   // {key, global contig, start, strand}
    uint2 *sorted;                 // hits bucketed by (key, strand): {global contig, start}
    uint32_t *bucketCount, *bucketOff, *bucketCursor;   // [2 * nKeys (+ 1)]
    uint32_t nKeys, hitCap;
    uint32_t *cnt;                 // [0] hits, [1] products, [2] probes that found an occupied slot
    const uint32_t *pairFwd, *pairRev;
    uint32_t nPairs, prodCap;
    int32_t badLo, badHi;          // params.disallowedLens = range(badLo, badHi + 1)
    uint32_t *removedAt;           // [nPairs] first outgroup genome with a disallowed product size, NONE = kept
    uint4 *products;               // {pair, global contig, size, 0}
};

__device__ __forceinline__ uint64_t rc64(uint64_t x) {   // reverse the 2-bit groups and complement (code ^ 1)
    uint64_t y = __brevll(x);
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
    return y ^ 0x5555555555555555ull;
}

__host__ __device__ __forceinline__ uint32_t key_hash(uint64_t w) {
    uint32_t h = (uint32_t)w * 0x9E3779B1u + (uint32_t)(w >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

__host__ __device__ __forceinline__ uint32_t key_tag(uint64_t w) {    // never 0
    uint32_t t = ((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x85EBCA77u)) * 0xC2B2AE3Du;
    t ^= t >> 16;
    return t | 1u;
}

// table slot: from the TAG, not from the hash the filter bit comes from -- a window that passes the filter has a hash
// next to a record's, so a slot taken from the same hash would always land in an occupied neighbourhood
__host__ __device__ __forceinline__ uint32_t key_slot(uint32_t tag, uint32_t size) {
    return (uint32_t)(((uint64_t)(tag * 0x9E3779B1u) * size) >> 32);
}

__device__ __forceinline__ uint32_t find_contig(const Args &p, uint32_t w) {
    uint32_t lo = 0, hi = p.nContigs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&p.cWordStart[mid]) <= w) lo = mid; else hi = mid;
    }
    return lo;
}

// ASCII -> 2 bits / base + invalid mask, one thread per 32-base word (str(contig.seq), removeOutgroupPrimers.py:221)
__global__ void __launch_bounds__(256) k_og_encode(const __grid_constant__ Args p) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= p.totalWords) return;
    const uint32_t c = find_contig(p, w);
    const uint32_t i = w - p.cWordStart[c], len = p.cLen[c];
    const uint32_t nvalid = min(32u, len - 32u * i);
    const uint8_t *src = p.raw + p.cRawStart[c] + 32ull * i;
    uintptr_t a = (uintptr_t)src;
    const uint32_t *src4 = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t x[9];
#pragma unroll
    for (int t = 0; t < 9; t++) x[t] = __ldg(src4 + t);   // the arena has 64 B of slack
    uint64_t packed = 0;
    uint32_t inval = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t v = sh ? __funnelshift_r(x[t], x[t + 1], sh) : x[t];
        const uint32_t vm = __vcmpeq4(v, 0x41414141u) | __vcmpeq4(v, 0x43434343u) | __vcmpeq4(v, 0x47474747u) | __vcmpeq4(v, 0x54545454u);
        const uint32_t tt = (v >> 1) & 0x03030303u;                                   // A0 C1 T2 G3
        uint32_t code = ((tt & 0x01010101u) << 1) | ((tt >> 1) & 0x01010101u);        // A0 T1 C2 G3
        code |= ~vm & 0x03030303u;
        const uint32_t pk = (code * 0x40100401u) >> 24;                               // byte 0 -> top 2 bits
        const uint32_t nb = ((~vm & 0x01010101u) * 0x08040201u) >> 24;                // byte 0 -> bit 3
        packed |= (uint64_t)pk << (56 - 8 * t);
        inval |= (nb & 0xFu) << (28 - 4 * t);
    }
    if (nvalid < 32) {
        inval |= 0xFFFFFFFFu >> nvalid;
        packed |= ~0ull >> (2 * nvalid);
    }
    p.seq2[w] = packed;
    p.nmask[w] = inval;
}

__device__ __forceinline__ void emit_hit(const Args &p, uint32_t key, uint32_t contig, uint32_t start, uint32_t strand) {
    if (key == NONE) return;
    const uint32_t n = atomicAdd(&p.cnt[0], 1u);
    atomicAdd(&p.bucketCount[2u * key + strand], 1u);
    if (n < p.hitCap) p.hits[n] = make_uint4(key, contig, start, strand);
}

// the table probe of one window whose filter bit is set (about one window in sixteen, plus the real hits)
__device__ __noinline__ void probe_window(const Args &p, int k, uint64_t h, uint64_t r, uint32_t hash, uint32_t c, uint32_t end,
                                          uint32_t &occupied) {
    const uint64_t canon = h < r ? h : r;
    const int ki = k - p.kmin;
    const uint32_t size = p.tabSize[ki];
    const uint2 *tab = p.tab + p.tabOff[ki];
    const uint32_t tag = key_tag(canon);
    uint32_t s = key_slot(tag, size);
    for (;;) {
        const uint2 e = __ldg(&tab[s]);
        if (e.x == 0u) return;                    // empty slot: no key of this length reads like the window
        occupied++;
        if (e.x == tag) {
            const uint4 rec = __ldg(&p.recs[e.y]);
            if (rec.x == (uint32_t)canon && rec.y == (uint32_t)(canon >> 32)) {
                // rec.z: the key whose string is the canonical word, rec.w: the key whose string is its rc
                const uint32_t plusKey = (h == canon) ? rec.z : rec.w;    // the key that reads like the window on (+)
                const uint32_t minusKey = (r == canon) ? rec.z : rec.w;   // the key that reads like the window's rc
                emit_hit(p, plusKey, c, end - (uint32_t)k + 1u, 0u);      // start = end - len + 1 (:51)
                emit_hit(p, minusKey, c, end, 1u);                        // len(seq) - start_in_rc - 1 (:54-55)
                return;
            }
        }
        if (++s == size) s = 0;
    }
}

// every window of every contig, both strands in one probe (replaces kmers.iter(fwd) / kmers.iter(rev), :49-60).
// One thread per 32-base word: the bases of the previous word give the windows that end in this one.  Per end
// position the U lengths of a group are hashed together and their filter words (1 bit per 1/16 record: a bitmap of
// 2 bytes per key that stays in L1 / L2) are loaded before any is looked at; only a set bit goes to the table.
// the windows that end in word w.  KMIN != 0: the lengths KMIN .. KMIN + NK - 1 are compile-time constants (constant
// shifts, filter geometry read from fixed parameter offsets; the reference's default 16..20 runs this way);
// KMIN == 0: lengths known at run time only, hashed NK at a time.  Per end position the filter words of all the
// lengths of a group are loaded before any is looked at; only a set bit goes to the table.
template <int KMIN, int NK>
__device__ __forceinline__ void scan_word(const Args &p, const uint32_t *bits, uint32_t w, uint32_t &occupied) {
    const uint32_t c = find_contig(p, w);
    const uint32_t i = w - __ldg(&p.cWordStart[c]);
    const uint64_t cur = p.seq2[w];
    const uint32_t cinv = p.nmask[w];
    uint64_t hf = i ? p.seq2[w - 1] : 0ull;          // the last 32 bases, newest lowest
    uint32_t inv = i ? p.nmask[w - 1] : 0xFFFFFFFFu; // their invalid bits, newest lowest (before a contig: all invalid)
    uint64_t rr = rc64(hf);                          // reverse complement of the last 32 bases, top aligned
    const int kmax = p.kmin + p.nk - 1;
    const uint32_t base0 = 32u * i;
#pragma unroll 2
    for (int t = 0; t < 32; t++) {
        const uint32_t code = (uint32_t)(cur >> (62 - 2 * t)) & 3u;
        hf = (hf << 2) | code;
        rr = (rr >> 2) | ((uint64_t)(code ^ 1u) << 62);
        inv = (inv << 1) | ((cinv >> (31 - t)) & 1u);
        if (inv & 1u) continue;                       // the newest base is in every window ending here
        const uint32_t end = base0 + t;               // (+) coordinate of the newest base
        if constexpr (KMIN != 0) {
            uint32_t hs[NK], fw[NK], fb[NK];
#pragma unroll
            for (int j = 0; j < NK; j++) {
                const int k = KMIN + j;
                const uint64_t h = hf & (~0ull >> (64 - 2 * k));
                const uint64_t r = rr >> (64 - 2 * k);
                hs[j] = key_hash(h < r ? h : r);
                fb[j] = __umulhi(hs[j], p.nBits[j]);
                fw[j] = (inv & (0xFFFFFFFFu >> (32 - k))) ? 0u : bits[p.bitOff[j] + (fb[j] >> 5)];
            }
#pragma unroll
            for (int j = 0; j < NK; j++) {
                if ((fw[j] >> (fb[j] & 31u)) & 1u) {
                    const int k = KMIN + j;
                    probe_window(p, k, hf & (~0ull >> (64 - 2 * k)), rr >> (64 - 2 * k), hs[j], c, end, occupied);
                }
            }
        } else {
            for (int kb = p.kmin; kb <= kmax; kb += NK) {
                uint32_t hs[NK], fw[NK], fb[NK];
#pragma unroll
                for (int j = 0; j < NK; j++) {
                    const int k = min(kb + j, 32), ki = min(kb + j, kmax) - p.kmin;
                    const bool live = kb + j <= kmax && !(inv & (0xFFFFFFFFu >> (32 - k)));
                    const uint64_t h = hf & (~0ull >> (64 - 2 * k));
                    const uint64_t r = rr >> (64 - 2 * k);
                    hs[j] = key_hash(h < r ? h : r);
                    fb[j] = __umulhi(hs[j], p.nBits[ki]);
                    fw[j] = live ? bits[p.bitOff[ki] + (fb[j] >> 5)] : 0u;
                }
#pragma unroll
                for (int j = 0; j < NK; j++) {
                    if ((fw[j] >> (fb[j] & 31u)) & 1u) {
                        const int k = kb + j;
                        probe_window(p, k, hf & (~0ull >> (64 - 2 * k)), rr >> (64 - 2 * k), hs[j], c, end, occupied);
                    }
                }
            }
        }
    }
}

// the filter in global memory (key sets too large for shared memory): one thread per word
template <int KMIN, int NK>
__global__ void __launch_bounds__(256) k_og_scan(const __grid_constant__ Args p) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= p.totalWords) return;
    uint32_t occupied = 0;
    scan_word<KMIN, NK>(p, p.bits, w, occupied);
    if (occupied) atomicAdd(&p.cnt[2], occupied);
}

// the table probe of one queued window: {forward word lo, hi, end, contig << 8 | k << 2}
__device__ __forceinline__ void probe_item(const Args &p, const uint4 it, uint32_t &occupied) {
    const int k = (int)(it.w >> 2) & 63, ki = k - p.kmin;
    const uint64_t h = (uint64_t)it.x | ((uint64_t)it.y << 32), r = rc64(h) >> (64 - 2 * k);
    const uint64_t canon = h < r ? h : r;
    const uint32_t size = p.tabSize[ki];
    const uint2 *tab = p.tab + p.tabOff[ki];
    const uint32_t tag = key_tag(canon);
    uint32_t s = key_slot(tag, size);
    for (;;) {
        const uint2 e = __ldg(&tab[s]);
        if (e.x == 0u) return;
        occupied++;
        if (e.x == tag) {
            const uint4 rec = __ldg(&p.recs[e.y]);
            if (rec.x == (uint32_t)canon && rec.y == (uint32_t)(canon >> 32)) {
                // rec.z: the key whose string is the canonical word, rec.w: the key whose string is its rc
                emit_hit(p, h == canon ? rec.z : rec.w, it.w >> 8, it.z - (uint32_t)k + 1u, 0u);   // start = end - len + 1 (:51)
                emit_hit(p, r == canon ? rec.z : rec.w, it.w >> 8, it.z, 1u);                      // len(seq) - start_in_rc - 1 (:54-55)
                return;
            }
        }
        if (++s == size) s = 0;
    }
}

constexpr uint32_t OGQ_ITEMS = 64;      // per-warp queue of filter-positive windows (16 B items)

// the filter of every length in shared memory (up to 192 KB: random LDS instead of one L1 line per lane and cycle);
// one persistent CTA per SM, every warp walks 32-word groups with a grid stride.  A window whose filter bit is set
// (about one in sixteen) is not probed at once -- a handful of lanes would wait for L2 while the others idle --
// but QUEUED per warp in shared memory; when 32 are waiting the warp probes them together, one per lane.
template <int KMIN, int NK>
__global__ void __launch_bounds__(1024, 1) k_og_scan_smem(const __grid_constant__ Args p) {
    extern __shared__ __align__(128) uint4 og_smem[];
    __shared__ __align__(8) uint64_t og_bar;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4 *q = og_smem + warp * OGQ_ITEMS;
    const uint32_t *bits = (const uint32_t *)(og_smem + 32u * OGQ_ITEMS);
    {   // the filter arrives as one bulk copy (cp.async.bulk -> UBLKCP) completing on an mbarrier
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&og_bar);
        if (threadIdx.x == 0) bulk::mbar_init(bar, 1u);
        __syncthreads();
        if (threadIdx.x == 0)
            bulk::load((uint32_t)__cvta_generic_to_shared(og_smem + 32u * OGQ_ITEMS), p.bits, ((p.bitWords + 3u) / 4u) * 16u, bar);
        bulk::wait(bar, 0u);
    }
    uint32_t occupied = 0, qn = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int kmin = KMIN ? KMIN : p.kmin, kmax = p.kmin + p.nk - 1;
    for (uint32_t wb = blockIdx.x * 1024u + warp * 32u; wb < p.totalWords; wb += gridDim.x * 1024u) {   // warp-uniform
        const uint32_t w = wb + lane;
        const bool valid = w < p.totalWords;
        uint32_t c = 0, i = 0, cinv = 0xFFFFFFFFu, inv = 0xFFFFFFFFu;
        uint64_t cur = 0, hf = 0;
        if (valid) {
            c = find_contig(p, w);
            i = w - __ldg(&p.cWordStart[c]);
            cur = p.seq2[w]; cinv = p.nmask[w];
            if (i) { hf = p.seq2[w - 1]; inv = p.nmask[w - 1]; }
        }
        uint64_t rr = rc64(hf);
        const uint32_t base0 = 32u * i, cw = c << 8;
        for (int t = 0; t < 32; t++) {
            const uint32_t code = (uint32_t)(cur >> (62 - 2 * t)) & 3u;
            hf = (hf << 2) | code;
            rr = (rr >> 2) | ((uint64_t)(code ^ 1u) << 62);
            inv = (inv << 1) | ((cinv >> (31 - t)) & 1u);
            const uint32_t end = base0 + t;
            for (int kb = kmin; kb <= kmax; kb += NK) {
                uint32_t fw[NK], fb[NK];
#pragma unroll
                for (int j = 0; j < NK; j++) {
                    const int k = KMIN ? KMIN + j : min(kb + j, 32), ki = KMIN ? j : min(kb + j, kmax) - p.kmin;
                    const bool live = (KMIN || kb + j <= kmax) && !(inv & (0xFFFFFFFFu >> (32 - k)));
                    const uint64_t h = hf & (~0ull >> (64 - 2 * k));
                    const uint64_t r = rr >> (64 - 2 * k);
                    fb[j] = __umulhi(key_hash(h < r ? h : r), p.nBits[ki]);
                    fw[j] = live ? bits[p.bitOff[ki] + (fb[j] >> 5)] : 0u;
                }
                // straight-line append: ballot, predicated store, uniform counter; the only branch is the (warp-uniform) drain
#pragma unroll
                for (int j = 0; j < NK; j++) {
                    const int k = KMIN ? KMIN + j : min(kb + j, 32);
                    const bool hit = (fw[j] >> (fb[j] & 31u)) & 1u;
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
                    if (hit) {
                        const uint64_t h = hf & (~0ull >> (64 - 2 * k));
                        q[qn + __popc(m & lt_mask)] = make_uint4((uint32_t)h, (uint32_t)(h >> 32), end, cw | ((uint32_t)k << 2));
                    }
                    qn += __popc(m);
                    if (qn >= 32u) {
                        __syncwarp();
                        qn -= 32u;
                        probe_item(p, q[qn + lane], occupied);
                        __syncwarp();
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane < qn) probe_item(p, q[lane], occupied);
    if (occupied) atomicAdd(&p.cnt[2], occupied);
}

// hits -> buckets of (key, strand) (the per-contig dicts of :215-226 turned inside out: contig is a column)
__global__ void __launch_bounds__(256) k_og_bucket(const __grid_constant__ Args p, uint32_t nHits) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nHits) return;
    const uint4 h = p.hits[j];
    const uint32_t b = 2u * h.x + h.w;
    const uint32_t dst = p.bucketOff[b] + atomicAdd(&p.bucketCursor[b], 1u);
    p.sorted[dst] = make_uint2(h.y, h.z);
}

// product sizes of one orientation of one pair: starts of `a` on (+) x starts of `b` on (-), same contig
// (__productSizesFromStartPositions, :65-88, called twice by __getOutgroupProductSizes, :91-122)
__device__ void pair_products(const Args &p, uint32_t pair, uint32_t a, uint32_t b) {
    const uint32_t a0 = p.bucketOff[2u * a], a1 = p.bucketOff[2u * a + 1u];
    const uint32_t b0 = p.bucketOff[2u * b + 1u], b1 = p.bucketOff[2u * b + 2u];
    for (uint32_t x = a0; x < a1; x++) {
        const uint2 f = p.sorted[x];
        for (uint32_t y = b0; y < b1; y++) {
            const uint2 r = p.sorted[y];
            if (r.x != f.x) continue;
            const int32_t len = (int32_t)r.y - (int32_t)f.y + 1;
            if (len <= 0) continue;                                   // facing away from each other (:83-85)
            if (len >= p.badLo && len <= p.badHi) {                   // pcrLen in params.disallowedLens (:283)
                atomicMin(&p.removedAt[pair], __ldg(&p.cGenome[f.x]));
            } else {
                const uint32_t n = atomicAdd(&p.cnt[1], 1u);
                if (n < p.prodCap) p.products[n] = make_uint4(pair, f.x, (uint32_t)len, 0u);
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_og_products(const __grid_constant__ Args p) {
    const uint32_t pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= p.nPairs) return;
    const uint32_t f = p.pairFwd[pair], r = p.pairRev[pair];
    pair_products(p, pair, f, r);     // fwd on (+), rev on (-)  (:107-108)
    pair_products(p, pair, r, f);     // rev on (+), fwd on (-)  (:116-117)
}

__global__ void __launch_bounds__(256) k_og_fill(uint32_t *a, uint32_t n, uint32_t v) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) a[j] = v;
}

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Genome {
    const uint8_t *bases;
    uint64_t n_bases;
    std::vector<uint64_t> off;
};

}  // namespace og

struct pfb_og {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    std::string err;
    std::vector<og::Genome> genomes;
    std::vector<uint64_t> keyWord;
    std::vector<uint8_t> keyK;
    std::vector<uint32_t> pairFwd, pairRev;
    int32_t badLo = 1, badHi = 0;
    int kmin = 0, nk = 0;
    og::Args a{};
    og::Buf raw, seq2, nmask, cWordStart, cLen, cGenome, cRawStart, tab, recs, bits, hits, sorted, bCount, bOff, bCursor, tiles, cnt,
        pFwd, pRev, removedAt, products;
    size_t hitCap = 0, prodCap = 0;
    bool keysUploaded = false, pairsUploaded = false, ran = false, filterInSmem = false;
    int scanMode = 0, numSMs = 148;      // scanMode 1: the filter stays in global memory (tests)
    uint32_t bitsPerRecord = 16;
    uint64_t tableBytes = 0;
    uint64_t nHits = 0, nProducts = 0, nBases = 0, nWindows = 0, nProbeOccupied = 0, nLaunches = 0, nReruns = 0;
    float msEncode = 0, msScan = 0, msBucket = 0, msProducts = 0, msTotal = 0, msUpload = 0;
    std::vector<uint32_t> cGenomeHost, cLocalHost;   // global contig -> genome, contig index within the genome
    std::vector<uint4> hitsHost, productsHost;
    std::vector<uint32_t> removedHost;
};

static std::string g_og_create_err;

namespace og {

int fail(pfb_og *o, int code, const std::string &msg) {
    if (o) o->err = msg; else g_og_create_err = msg;
    return code;
}

#define OGCK(call)                                                                                     \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return og::fail(o, PFB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int ensure(pfb_og *o, Buf &b, size_t bytes) {
    if (b.p && bytes <= b.cap) return PFB_OK;
    if (b.p) OGCK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    OGCK(cudaMalloc(&b.p, want));
    b.cap = want;
    return PFB_OK;
}

void release(Buf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

inline uint64_t rc_word(uint64_t w, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | ((w & 3u) ^ 1u); w >>= 2; }
    return r;
}

// the key tables: one per length, canonical word -> (key reading like it, key reading like its rc)
int upload_keys(pfb_og *o) {
    const size_t n = o->keyWord.size();
    int kmin = 33, kmax = 0;
    for (size_t i = 0; i < n; i++) { kmin = std::min<int>(kmin, o->keyK[i]); kmax = std::max<int>(kmax, o->keyK[i]); }
    if (n == 0) { kmin = 1; kmax = 1; }
    o->kmin = kmin; o->nk = kmax - kmin + 1;
    std::vector<uint32_t> perK(o->nk, 0);
    for (size_t i = 0; i < n; i++) perK[o->keyK[i] - kmin]++;
    Args &a = o->a;
    // the canonical records: a key and the key that is its reverse complement share one
    std::vector<uint4> recs;
    std::vector<uint8_t> recK;
    {
        std::vector<std::pair<std::pair<int, uint64_t>, uint32_t>> order(n);     // ((k, canonical word), key)
        for (size_t i = 0; i < n; i++)
            order[i] = {{o->keyK[i], std::min(o->keyWord[i], rc_word(o->keyWord[i], o->keyK[i]))}, (uint32_t)i};
        std::sort(order.begin(), order.end());
        for (size_t j = 0; j < n; j++) {
            const uint32_t i = order[j].second;
            const int k = order[j].first.first;
            const uint64_t canon = order[j].first.second, w = o->keyWord[i];
            if (j == 0 || order[j - 1].first != order[j].first) {
                recs.push_back(make_uint4((uint32_t)canon, (uint32_t)(canon >> 32), NONE, NONE));
                recK.push_back((uint8_t)k);
            }
            uint4 &e = recs.back();
            if (w == canon) { if (e.z != NONE) return fail(o, PFB_EINVAL, "duplicate key"); e.z = i; }
            if (rc_word(w, k) == canon) {     // (a palindrome fills both: it reads the same on both strands)
                if (e.w != NONE && e.w != i) return fail(o, PFB_EINVAL, "duplicate key");
                e.w = i;
            }
        }
    }
    std::fill(perK.begin(), perK.end(), 0u);
    for (uint8_t k : recK) perK[k - kmin]++;
    uint32_t total = 0;
    for (int ki = 0; ki < o->nk; ki++) {
        const uint32_t sz = std::max(64u, 3u * perK[ki]);
        a.tabOff[ki] = total; a.tabSize[ki] = sz;
        total += sz;
    }
    a.tabOff[o->nk] = total;
    std::vector<uint2> tab(total, make_uint2(0u, 0u));
    for (size_t j = 0; j < recs.size(); j++) {
        const int ki = recK[j] - kmin;
        const uint64_t canon = (uint64_t)recs[j].x | ((uint64_t)recs[j].y << 32);
        uint2 *t = tab.data() + a.tabOff[ki];
        uint32_t s = key_slot(key_tag(canon), a.tabSize[ki]);
        while (t[s].x != 0u) if (++s == a.tabSize[ki]) s = 0;
        t[s] = make_uint2(key_tag(canon), (uint32_t)j);
    }
    // the filter: 16 bits per record while that fits the shared-memory budget, down to 4 bits per record (one
    // window in four passes); beyond that it stays in global memory at 16 bits per record
    uint32_t perRec = 16;
    const uint64_t budget = (uint64_t)SMEM_FILTER_BYTES * 8u - 1024ull * 33ull - 128ull * 33ull;
    while (perRec > 4 && (uint64_t)perRec * recs.size() > budget) perRec >>= 1;
    o->filterInSmem = (uint64_t)perRec * recs.size() <= budget && o->scanMode != 1;
    if (!o->filterInSmem) perRec = 16;
    uint32_t bitWords = 0;
    for (int ki = 0; ki < o->nk; ki++) {
        a.nBits[ki] = std::max(1024u, perRec * perK[ki]);
        a.bitOff[ki] = bitWords;
        bitWords += (a.nBits[ki] + 127u) / 128u * 4u;       // 16-byte granules (the copy to shared memory)
    }
    a.bitWords = bitWords;
    o->bitsPerRecord = perRec;
    std::vector<uint32_t> bits(bitWords, 0u);
    for (size_t j = 0; j < recs.size(); j++) {
        const int ki = recK[j] - kmin;
        const uint64_t canon = (uint64_t)recs[j].x | ((uint64_t)recs[j].y << 32);
        const uint32_t b = (uint32_t)(((uint64_t)key_hash(canon) * a.nBits[ki]) >> 32);
        bits[a.bitOff[ki] + (b >> 5)] |= 1u << (b & 31u);
    }
    int rc;
    if ((rc = ensure(o, o->bits, (size_t)bitWords * 4)) != PFB_OK) return rc;
    OGCK(cudaMemcpyAsync(o->bits.p, bits.data(), (size_t)bitWords * 4, cudaMemcpyHostToDevice, o->stream));
    a.bits = (const uint32_t *)o->bits.p;
    if ((rc = ensure(o, o->tab, (size_t)total * 8)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->recs, recs.size() * 16 + 16)) != PFB_OK) return rc;
    OGCK(cudaMemcpyAsync(o->tab.p, tab.data(), (size_t)total * 8, cudaMemcpyHostToDevice, o->stream));
    if (!recs.empty()) OGCK(cudaMemcpyAsync(o->recs.p, recs.data(), recs.size() * 16, cudaMemcpyHostToDevice, o->stream));
    OGCK(cudaStreamSynchronize(o->stream));
    o->tableBytes = (uint64_t)total * 8 + recs.size() * 16 + (uint64_t)bitWords * 4;
    if ((rc = ensure(o, o->bCount, (2 * n + 2) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->bOff, (2 * n + 2) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->bCursor, (2 * n + 2) * 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->tiles, ((2 * n + 2) / devscan::TILE + 4) * 4)) != PFB_OK) return rc;
    a.tab = (const uint2 *)o->tab.p; a.recs = (const uint4 *)o->recs.p;
    a.kmin = kmin; a.nk = o->nk; a.nKeys = (uint32_t)n;
    a.bucketCount = (uint32_t *)o->bCount.p; a.bucketOff = (uint32_t *)o->bOff.p; a.bucketCursor = (uint32_t *)o->bCursor.p;
    o->keysUploaded = true;
    return PFB_OK;
}

int upload_pairs(pfb_og *o) {
    const size_t n = o->pairFwd.size();
    int rc;
    if ((rc = ensure(o, o->pFwd, n * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->pRev, n * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->removedAt, n * 4 + 4)) != PFB_OK) return rc;
    if (n) {
        OGCK(cudaMemcpyAsync(o->pFwd.p, o->pairFwd.data(), n * 4, cudaMemcpyHostToDevice, o->stream));
        OGCK(cudaMemcpyAsync(o->pRev.p, o->pairRev.data(), n * 4, cudaMemcpyHostToDevice, o->stream));
        OGCK(cudaStreamSynchronize(o->stream));
    }
    Args &a = o->a;
    a.pairFwd = (const uint32_t *)o->pFwd.p; a.pairRev = (const uint32_t *)o->pRev.p; a.nPairs = (uint32_t)n;
    a.removedAt = (uint32_t *)o->removedAt.p;
    a.badLo = o->badLo; a.badHi = o->badHi;
    o->pairsUploaded = true;
    return PFB_OK;
}

// contig tables + the ASCII bases (pinned host memory makes the copies asynchronous)
int upload_genomes(pfb_og *o) {
    std::vector<uint32_t> cWordStart, cLen, cGenome;
    std::vector<uint64_t> cRawStart;
    o->cLocalHost.clear();
    uint64_t words = 0, rawBytes = 0, windows = 0;
    for (size_t g = 0; g < o->genomes.size(); g++) {
        const Genome &G = o->genomes[g];
        for (size_t c = 0; c + 1 < G.off.size(); c++) {
            const uint64_t len = G.off[c + 1] - G.off[c];
            if (len == 0) continue;                       // an empty record has no windows
            if (len >= (1ull << 31)) return fail(o, PFB_ETOOLARGE, "an outgroup contig exceeds 2^31 bases");
            cWordStart.push_back((uint32_t)words); cLen.push_back((uint32_t)len); cGenome.push_back((uint32_t)g);
            cRawStart.push_back(rawBytes + G.off[c]);
            o->cLocalHost.push_back((uint32_t)c);
            words += (len + 31) / 32;
            for (int k = o->kmin; k < o->kmin + o->nk; k++) if (len >= (uint64_t)k) windows += len - k + 1;
        }
        rawBytes += G.n_bases;
    }
    if (words >= (1ull << 32)) return fail(o, PFB_ETOOLARGE, "the outgroup genomes exceed 2^37 bases in one run");
    if (cWordStart.size() >= (1u << 24)) return fail(o, PFB_ETOOLARGE, "more than 2^24 outgroup contigs in one run");
    o->cGenomeHost = cGenome;
    o->nBases = rawBytes; o->nWindows = windows;
    const size_t nc = cWordStart.size();
    int rc;
    if ((rc = ensure(o, o->raw, rawBytes + 64)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->seq2, words * 8 + 8)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->nmask, words * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->cWordStart, nc * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->cLen, nc * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->cGenome, nc * 4 + 4)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->cRawStart, nc * 8 + 8)) != PFB_OK) return rc;
    OGCK(cudaEventRecord(o->ev[0], o->stream));
    uint64_t at = 0;
    for (const Genome &G : o->genomes) {
        if (G.n_bases) OGCK(cudaMemcpyAsync((uint8_t *)o->raw.p + at, G.bases, G.n_bases, cudaMemcpyHostToDevice, o->stream));
        at += G.n_bases;
    }
    if (nc) {
        OGCK(cudaMemcpyAsync(o->cWordStart.p, cWordStart.data(), nc * 4, cudaMemcpyHostToDevice, o->stream));
        OGCK(cudaMemcpyAsync(o->cLen.p, cLen.data(), nc * 4, cudaMemcpyHostToDevice, o->stream));
        OGCK(cudaMemcpyAsync(o->cGenome.p, cGenome.data(), nc * 4, cudaMemcpyHostToDevice, o->stream));
        OGCK(cudaMemcpyAsync(o->cRawStart.p, cRawStart.data(), nc * 8, cudaMemcpyHostToDevice, o->stream));
    }
    OGCK(cudaEventRecord(o->ev[1], o->stream));
    OGCK(cudaStreamSynchronize(o->stream));      // the tables above are host temporaries
    Args &a = o->a;
    a.raw = (const uint8_t *)o->raw.p; a.seq2 = (uint64_t *)o->seq2.p; a.nmask = (uint32_t *)o->nmask.p;
    a.cWordStart = (const uint32_t *)o->cWordStart.p; a.cLen = (const uint32_t *)o->cLen.p;
    a.cGenome = (const uint32_t *)o->cGenome.p; a.cRawStart = (const uint64_t *)o->cRawStart.p;
    a.nContigs = (uint32_t)nc; a.totalWords = (uint32_t)words;
    return PFB_OK;
}

int compute(pfb_og *o) {
    Args &a = o->a;
    int rc;
    const uint32_t nB = 2u * a.nKeys;
    if (o->hitCap == 0) o->hitCap = std::max<size_t>(1u << 16, 8ull * a.nKeys);
    if (o->prodCap == 0) o->prodCap = std::max<size_t>(1u << 16, 2ull * a.nPairs);
    if ((rc = ensure(o, o->cnt, 64)) != PFB_OK) return rc;
    a.cnt = (uint32_t *)o->cnt.p;
    o->nLaunches = 0;
    for (int attempt = 0;; attempt++) {
        if ((rc = ensure(o, o->hits, o->hitCap * 16)) != PFB_OK) return rc;
        if ((rc = ensure(o, o->sorted, o->hitCap * 8)) != PFB_OK) return rc;
        if ((rc = ensure(o, o->products, o->prodCap * 16)) != PFB_OK) return rc;
        a.hits = (uint4 *)o->hits.p; a.sorted = (uint2 *)o->sorted.p; a.products = (uint4 *)o->products.p;
        a.hitCap = (uint32_t)o->hitCap; a.prodCap = (uint32_t)o->prodCap;
        OGCK(cudaEventRecord(o->ev[2], o->stream));
        OGCK(cudaMemsetAsync(o->cnt.p, 0, 64, o->stream));
        OGCK(cudaMemsetAsync(o->bCount.p, 0, (size_t)(nB + 2) * 4, o->stream));
        OGCK(cudaMemsetAsync(o->bCursor.p, 0, (size_t)(nB + 2) * 4, o->stream));
        if (a.nPairs) { k_og_fill<<<(a.nPairs + 255) / 256, 256, 0, o->stream>>>(a.removedAt, a.nPairs, NONE); o->nLaunches++; }
        if (a.totalWords) {
            k_og_encode<<<(a.totalWords + 255) / 256, 256, 0, o->stream>>>(a);
            OGCK(cudaEventRecord(o->ev[3], o->stream));
            const bool def = a.kmin == 16 && a.nk == 5;       // params.minLen / maxLen defaults (bin/Parameters.py:34-35)
            if (o->filterInSmem) {
                const size_t smem = ((size_t)a.bitWords * 4 + 15) / 16 * 16 + 32u * OGQ_ITEMS * 16u;
                if (def) k_og_scan_smem<16, 5><<<o->numSMs, 1024, smem, o->stream>>>(a);
                else k_og_scan_smem<0, 4><<<o->numSMs, 1024, smem, o->stream>>>(a);
            } else {
                if (def) k_og_scan<16, 5><<<(a.totalWords + 255) / 256, 256, 0, o->stream>>>(a);
                else k_og_scan<0, 4><<<(a.totalWords + 255) / 256, 256, 0, o->stream>>>(a);
            }
            o->nLaunches += 2;
        } else {
            OGCK(cudaEventRecord(o->ev[3], o->stream));
        }
        OGCK(cudaEventRecord(o->ev[4], o->stream));
        uint32_t cnt[4] = {0, 0, 0, 0};
        OGCK(cudaMemcpyAsync(cnt, o->cnt.p, 16, cudaMemcpyDeviceToHost, o->stream));
        OGCK(cudaStreamSynchronize(o->stream));
        if (cnt[0] > o->hitCap) { o->hitCap = (size_t)cnt[0] + cnt[0] / 4; o->nReruns++; continue; }
        o->nHits = cnt[0]; o->nProbeOccupied = cnt[2];
        // bucket the hits by (key, strand), then one thread per pair
        devscan::exscan(a.bucketCount, nB, a.bucketOff, (uint32_t *)o->tiles.p, a.cnt + 3, o->stream);
        o->nLaunches += 3;
        if (cnt[0]) { k_og_bucket<<<(cnt[0] + 255) / 256, 256, 0, o->stream>>>(a, cnt[0]); o->nLaunches++; }
        OGCK(cudaEventRecord(o->ev[5], o->stream));
        if (a.nPairs) { k_og_products<<<(a.nPairs + 127) / 128, 128, 0, o->stream>>>(a); o->nLaunches++; }
        OGCK(cudaEventRecord(o->ev[0], o->stream));
        OGCK(cudaMemcpyAsync(cnt, o->cnt.p, 16, cudaMemcpyDeviceToHost, o->stream));
        OGCK(cudaStreamSynchronize(o->stream));
        OGCK(cudaGetLastError());
        if (cnt[1] > o->prodCap) { o->prodCap = (size_t)cnt[1] + cnt[1] / 4; o->nReruns++; continue; }
        o->nProducts = cnt[1];
        break;
    }
    OGCK(cudaEventElapsedTime(&o->msEncode, o->ev[2], o->ev[3]));
    OGCK(cudaEventElapsedTime(&o->msScan, o->ev[3], o->ev[4]));
    OGCK(cudaEventElapsedTime(&o->msBucket, o->ev[4], o->ev[5]));
    OGCK(cudaEventElapsedTime(&o->msProducts, o->ev[5], o->ev[0]));
    o->msTotal = o->msEncode + o->msScan + o->msBucket + o->msProducts;
    // results to the host
    o->hitsHost.resize(o->nHits); o->productsHost.resize(o->nProducts); o->removedHost.resize(a.nPairs);
    if (o->nHits) OGCK(cudaMemcpyAsync(o->hitsHost.data(), o->hits.p, o->nHits * 16, cudaMemcpyDeviceToHost, o->stream));
    if (o->nProducts) OGCK(cudaMemcpyAsync(o->productsHost.data(), o->products.p, o->nProducts * 16, cudaMemcpyDeviceToHost, o->stream));
    if (a.nPairs) OGCK(cudaMemcpyAsync(o->removedHost.data(), o->removedAt.p, (size_t)a.nPairs * 4, cudaMemcpyDeviceToHost, o->stream));
    OGCK(cudaStreamSynchronize(o->stream));
    o->ran = true;
    return PFB_OK;
}

}  // namespace og

extern "C" {

int pfb_og_create(pfb_og **out, int device) {
    if (!out) return PFB_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return og::fail(nullptr, PFB_ECUDA, "no CUDA device (primerforge_b200 has no CPU path)");
    if (device < 0 || device >= n) return og::fail(nullptr, PFB_EINVAL, "no such device");
    pfb_og *ctx = new pfb_og();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return og::fail(nullptr, PFB_ECUDA, "cudaSetDevice / cudaStreamCreate failed");
    }
    cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device);
    if (cudaFuncSetAttribute(og::k_og_scan_smem<16, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(og::SMEM_FILTER_BYTES + 32u * og::OGQ_ITEMS * 16u)) != cudaSuccess ||
        cudaFuncSetAttribute(og::k_og_scan_smem<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(og::SMEM_FILTER_BYTES + 32u * og::OGQ_ITEMS * 16u)) != cudaSuccess) {
        delete ctx;
        return og::fail(nullptr, PFB_ECUDA, "this device cannot give a CTA 192 KB of shared memory");
    }
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return og::fail(nullptr, PFB_ECUDA, "cudaEventCreate failed"); }
    *out = ctx;
    return PFB_OK;
}

void pfb_og_destroy(pfb_og *o) {
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->stream) cudaStreamSynchronize(o->stream);
    for (og::Buf *b : {&o->raw, &o->seq2, &o->nmask, &o->cWordStart, &o->cLen, &o->cGenome, &o->cRawStart, &o->tab, &o->recs, &o->bits, &o->hits,
                       &o->sorted, &o->bCount, &o->bOff, &o->bCursor, &o->tiles, &o->cnt, &o->pFwd, &o->pRev, &o->removedAt,
                       &o->products})
        og::release(*b);
    for (auto &e : o->ev) if (e) cudaEventDestroy(e);
    if (o->stream) cudaStreamDestroy(o->stream);
    delete o;
}

const char *pfb_og_last_error(const pfb_og *o) { return o ? o->err.c_str() : g_og_create_err.c_str(); }

int pfb_og_set_keys(pfb_og *o, const uint64_t *word, const uint8_t *k, uint32_t n_keys) {
    if (!o) return PFB_EINVAL;
    if (n_keys && (!word || !k)) return og::fail(o, PFB_EINVAL, "bad key arguments");
    if (n_keys >= (1u << 30)) return og::fail(o, PFB_ETOOLARGE, "too many keys");
    for (uint32_t i = 0; i < n_keys; i++) {
        if (k[i] < 1 || k[i] > 32) return og::fail(o, PFB_EINVAL, "key length must be 1..32");
        if (k[i] < 32 && (word[i] >> (2 * k[i]))) return og::fail(o, PFB_EINVAL, "key word has bits beyond its length");
    }
    o->keyWord.assign(word, word + n_keys);
    o->keyK.assign(k, k + n_keys);
    o->pairFwd.clear(); o->pairRev.clear();
    o->keysUploaded = false; o->pairsUploaded = false; o->ran = false; o->hitCap = 0;
    return PFB_OK;
}

int pfb_og_set_pairs(pfb_og *o, const uint32_t *fwd_key, const uint32_t *rev_key, uint32_t n_pairs, int32_t disallowed_lo,
                     int32_t disallowed_hi) {
    if (!o) return PFB_EINVAL;
    if (n_pairs && (!fwd_key || !rev_key)) return og::fail(o, PFB_EINVAL, "bad pair arguments");
    for (uint32_t i = 0; i < n_pairs; i++)
        if (fwd_key[i] >= o->keyWord.size() || rev_key[i] >= o->keyWord.size()) return og::fail(o, PFB_EINVAL, "pair refers to a key that was not set");
    o->pairFwd.assign(fwd_key, fwd_key + n_pairs);
    o->pairRev.assign(rev_key, rev_key + n_pairs);
    o->badLo = disallowed_lo; o->badHi = disallowed_hi;
    o->pairsUploaded = false; o->ran = false; o->prodCap = 0;
    return PFB_OK;
}

int pfb_og_add_genome(pfb_og *o, const uint8_t *bases, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs) {
    if (!o) return PFB_EINVAL;
    if ((!bases && n_bases) || !contig_off) return og::fail(o, PFB_EINVAL, "bad genome arguments");
    if (contig_off[0] != 0 || contig_off[n_contigs] != n_bases) return og::fail(o, PFB_EINVAL, "contig offsets must span [0, n_bases]");
    for (uint32_t c = 0; c < n_contigs; c++)
        if (contig_off[c + 1] < contig_off[c]) return og::fail(o, PFB_EINVAL, "contig offsets must be non-decreasing");
    og::Genome g;
    g.bases = bases; g.n_bases = n_bases;
    g.off.assign(contig_off, contig_off + n_contigs + 1);
    o->genomes.push_back(std::move(g));
    o->ran = false;
    return PFB_OK;
}

int pfb_og_set_scan_mode(pfb_og *o, int mode) {
    if (!o || mode < 0 || mode > 1) return PFB_EINVAL;
    o->scanMode = mode;
    o->keysUploaded = false; o->ran = false;
    return PFB_OK;
}

int pfb_og_clear_genomes(pfb_og *o) {
    if (!o) return PFB_EINVAL;
    o->genomes.clear();
    o->ran = false;
    return PFB_OK;
}

int pfb_og_run(pfb_og *o) {
    if (!o) return PFB_EINVAL;
    if (cudaSetDevice(o->device) != cudaSuccess) return og::fail(o, PFB_ECUDA, "cudaSetDevice failed");
    int rc;
    if (!o->keysUploaded && (rc = og::upload_keys(o)) != PFB_OK) return rc;
    if (!o->pairsUploaded && (rc = og::upload_pairs(o)) != PFB_OK) return rc;
    if ((rc = og::upload_genomes(o)) != PFB_OK) return rc;
    if (cudaEventElapsedTime(&o->msUpload, o->ev[0], o->ev[1]) != cudaSuccess) o->msUpload = 0;
    return og::compute(o);
}

int pfb_og_compute(pfb_og *o) {
    if (!o) return PFB_EINVAL;
    if (!o->ran) return og::fail(o, PFB_ESTATE, "pfb_og_compute repeats the last pfb_og_run on the resident genomes");
    if (cudaSetDevice(o->device) != cudaSuccess) return og::fail(o, PFB_ECUDA, "cudaSetDevice failed");
    return og::compute(o);
}

int pfb_og_counts(pfb_og *o, uint64_t *n_hits, uint64_t *n_products, uint64_t *n_removed) {
    if (!o) return PFB_EINVAL;
    if (!o->ran) return og::fail(o, PFB_ESTATE, "no run yet");
    if (n_hits) *n_hits = o->nHits;
    if (n_products) *n_products = o->nProducts;
    if (n_removed) {
        uint64_t r = 0;
        for (uint32_t v : o->removedHost) r += v != og::NONE;
        *n_removed = r;
    }
    return PFB_OK;
}

int pfb_og_hits(pfb_og *o, pfb_og_hit *out, uint64_t cap) {
    if (!o || (!out && cap)) return PFB_EINVAL;
    if (!o->ran) return og::fail(o, PFB_ESTATE, "no run yet");
    if (cap < o->nHits) return og::fail(o, PFB_EINVAL, "hit buffer too small");
    for (uint64_t i = 0; i < o->nHits; i++) {
        const uint4 h = o->hitsHost[i];
        out[i].key = h.x; out[i].genome = o->cGenomeHost[h.y]; out[i].contig = o->cLocalHost[h.y]; out[i].start = h.z; out[i].strand = h.w;
    }
    return PFB_OK;
}

int pfb_og_products(pfb_og *o, pfb_og_product *out, uint64_t cap) {
    if (!o || (!out && cap)) return PFB_EINVAL;
    if (!o->ran) return og::fail(o, PFB_ESTATE, "no run yet");
    if (cap < o->nProducts) return og::fail(o, PFB_EINVAL, "product buffer too small");
    for (uint64_t i = 0; i < o->nProducts; i++) {
        const uint4 h = o->productsHost[i];
        out[i].pair = h.x; out[i].genome = o->cGenomeHost[h.y]; out[i].contig = o->cLocalHost[h.y]; out[i].size = (int32_t)h.z;
    }
    return PFB_OK;
}

int pfb_og_removed_at(pfb_og *o, uint32_t *out, uint64_t cap) {
    if (!o || (!out && cap)) return PFB_EINVAL;
    if (!o->ran) return og::fail(o, PFB_ESTATE, "no run yet");
    if (cap < o->removedHost.size()) return og::fail(o, PFB_EINVAL, "buffer too small");
    if (!o->removedHost.empty()) memcpy(out, o->removedHost.data(), o->removedHost.size() * 4);
    return PFB_OK;
}

int pfb_og_timings(pfb_og *o, pfb_og_timing *out) {
    if (!o || !out) return PFB_EINVAL;
    memset(out, 0, sizeof(*out));
    out->ms_upload = o->msUpload; out->ms_encode = o->msEncode; out->ms_scan = o->msScan; out->ms_bucket = o->msBucket;
    out->ms_products = o->msProducts; out->ms_total = o->msTotal;
    out->n_bases = o->nBases; out->n_windows = o->nWindows; out->n_hits = o->nHits; out->n_products = o->nProducts;
    out->n_probe_occupied = o->nProbeOccupied; out->n_launches = o->nLaunches; out->n_reruns = o->nReruns;
    out->table_bytes = o->tableBytes;
    out->filter_in_smem = o->filterInSmem ? 1u : 0u; out->filter_bits_per_record = o->bitsPerRecord;
    return PFB_OK;
}

}  // extern "C"
