// pfb_pairs.cu -- B200 (sm_100a) pair search of primerForge, from scratch.
//
// What the reference computes (bin/getPrimerPairs.py:490-564; file:line citations are into /root/reference): the
// candidates of every genome are binned by a sequential walk over their sorted lists (:39-98: a candidate joins the
// current bin while it overlaps its predecessor and the bin stays within maxBinSize); in the first genome's bins the
// primers that are substrings of a longer primer of the same bin are dropped (:10-36, with Python's
// edit-the-list-you-iterate semantics); all pairs of bins of a contig that can give an allowed product size are
// listed (:101-143) and each contributes the FIRST (forward, reverse) primer pair -- in bin order -- with G/C 3' ends,
// an allowed product size and a Tm difference within bounds (:283-341; the heterodimer check of :330-341 is primer3's
// and stays on the host); every pair is then looked up in every other genome: same contig, same strand, allowed
// product size (:367-464), and gets the bin numbers of that genome (:467-502).
//
// How it is computed here.  The candidates are columns (word, length, Tm; per genome the list order with contig,
// leftmost coordinate, strand).  Binning: every list position computes J = "where would the next bin start if a bin
// started here" independently (k_pp_jump); the bin starts are the orbit of J from the start of the list, walked by
// one CTA per genome through shared-memory tiles of J (k_pp_orbit); a device-wide scan turns the start flags into bin
// numbers.  The substring reduction is one thread per bin working on the bin's members in place (the list-iterator
// quirk becomes a "skip the next one" flag), bin pairs + the first suitable primer pair are one thread per first bin
// in two passes (count, scan, emit: the pairs come out in the reference's order without a sort), the cross-genome
// look-up is one thread per (pair, genome).

#include "../../include/pfb200.h"
#include "pfb_devscan.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace pp {

struct Args {
    uint32_t S, G;                 // candidates, genomes
    const uint64_t *word;          // [S] the k-mer as it reads on the first genome's (+) strand
    const uint8_t *k;              // [S]
    const double *tm;              // [S]
    // per genome, list order (position p): candidate id, contig (dense, non-decreasing), leftmost coordinate, strand
    const uint32_t *gid, *gcontig, *gleft;      // [G][S]
    const uint8_t *gstrand;                     // [G][S]
    // per genome, by candidate id
    uint32_t *ctOf, *lfOf, *binOf;              // [G][S]
    uint8_t *stOf;                              // [G][S]
    uint32_t *seen;                             // [G][S] how often a genome lists a candidate (must be 1)
    uint32_t *jump;                // [G][S] J
    uint32_t *isStart;             // [G][S + 1] 1 = a bin starts at this list position
    uint32_t *binIdx;              // [G][S + 1] exclusive scan of isStart (bins before this position)
    uint32_t *contigFirst;         // [G][S] list position where the contig of position p begins
    // first genome
    uint32_t *binPos;              // [nb + 1] list position of every bin start (+ S)
    uint8_t *alive;                // [S] by list position: 0 = removed by __reduceBinSize
    uint32_t *binFirst, *binLast;  // [nb] first / last alive list position of every bin
    uint32_t *pairCount, *pairOff; // [nb + 1] pairs contributed by every first bin; exclusive scan
    uint4 *pairsA;                 // [nPairs] {fwd id, rev id, contig, pcrLen}
    uint2 *pairsB;                 // [nPairs] {fbin, rbin}
    uint32_t *shared;              // [G][nPairs] product length | 1 << 31 when the pair has a Product in that genome
    uint8_t *kept;                 // [nPairs] 1 = a Product in every genome
    uint32_t *cnt;                 // [0] bins of the first genome, [1] pairs, [2] error flags, [4] kept pairs
    uint32_t nb, pairCap;
    int32_t maxBin, minPrimerLen, minPcr, maxPcr;
    double maxTmDiff;
};

enum { ERR_DUP = 1u };

// candidate id -> (contig, leftmost, strand) per genome (the dict kmers[name] of :349-362, 385-390)
__global__ void __launch_bounds__(256) k_pp_scatter(const __grid_constant__ Args a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (p >= a.S) return;
    const size_t o = (size_t)g * a.S;
    const uint32_t id = a.gid[o + p];
    if (id >= a.S) { atomicOr(&a.cnt[2], ERR_DUP); return; }
    if (atomicAdd(&a.seen[o + id], 1u)) atomicOr(&a.cnt[2], ERR_DUP);
    a.ctOf[o + id] = a.gcontig[o + p];
    a.lfOf[o + id] = a.gleft[o + p];
    a.stOf[o + id] = a.gstrand[o + p];
}

// contigFirst[p]: the list position where the contig of position p begins (contig numbers are non-decreasing along the list)
__global__ void __launch_bounds__(256) k_pp_contig_first(const __grid_constant__ Args a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (p >= a.S) return;
    const size_t o = (size_t)g * a.S;
    const uint32_t c = a.gcontig[o + p];
    uint32_t lo = 0, hi = p;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.gcontig[o + mid] < c) lo = mid + 1; else hi = mid; }
    a.contigFirst[o + p] = lo;
}

// J[p]: the list position at which the next bin starts if a bin starts at p (__binCandidateKmers :66-90 with
// binStart = start[p]): the run of followers that overlap their predecessor (curStart < prevEnd) within maxBinSize
__global__ void __launch_bounds__(256) k_pp_jump(const __grid_constant__ Args a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (p >= a.S) return;
    const size_t o = (size_t)g * a.S;
    const uint32_t c = a.gcontig[o + p];
    const int64_t binStart = a.gleft[o + p];
    int64_t prevEnd = binStart + a.k[a.gid[o + p]] - 1;
    uint32_t j = p + 1;
    while (j < a.S && a.gcontig[o + j] == c) {
        const int64_t curStart = a.gleft[o + j], curEnd = curStart + a.k[a.gid[o + j]] - 1;
        if (!(curStart < prevEnd && curEnd - binStart <= a.maxBin)) break;      // :80
        prevEnd = curEnd;
        j++;
    }
    a.jump[o + p] = j;
}

// the bin starts = the orbit of J from position 0; one CTA per genome, J staged through shared memory in tiles.  Per tile
// all threads build the powers J^2, J^4 .. J^32 (saturating at the tile end), ONE thread walks the tile in steps of J^32
// (about ten dependent shared-memory loads instead of a few hundred), then all threads fill in the steps between its
// anchors: for r = 4 .. 0 every marked position marks J^(2^r) of itself.
constexpr uint32_t ORBIT_TILE = 4096, ORBIT_LEVELS = 6;
constexpr uint32_t ORBIT_SMEM = ORBIT_LEVELS * ORBIT_TILE * 4u + ORBIT_TILE;
__global__ void __launch_bounds__(1024) k_pp_orbit(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) uint32_t orbit_smem[];
    uint32_t *lv = orbit_smem;                                          // [LEVELS][TILE] absolute positions
    uint8_t *sF = (uint8_t *)(orbit_smem + ORBIT_LEVELS * ORBIT_TILE);  // [TILE] 1 = a bin starts here
    __shared__ uint32_t sCur;
    const uint32_t g = blockIdx.x;
    const size_t o = (size_t)g * a.S, of = (size_t)g * (a.S + 1);
    if (threadIdx.x == 0) sCur = 0;
    for (uint32_t t0 = 0; t0 < a.S; t0 += ORBIT_TILE) {
        const uint32_t n = min(ORBIT_TILE, a.S - t0), end = t0 + n;
        for (uint32_t j = threadIdx.x; j < n; j += 1024) { lv[j] = a.jump[o + t0 + j]; sF[j] = 0; }
        __syncthreads();
        for (uint32_t L = 1; L < ORBIT_LEVELS; L++) {
            const uint32_t *src = lv + (L - 1) * ORBIT_TILE;
            uint32_t *dst = lv + L * ORBIT_TILE;
            for (uint32_t j = threadIdx.x; j < n; j += 1024) { const uint32_t x = src[j]; dst[j] = x < end ? src[x - t0] : x; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const uint32_t *top = lv + (ORBIT_LEVELS - 1) * ORBIT_TILE;
            uint32_t cur = sCur;
            while (cur < end) { sF[cur - t0] = 1; cur = top[cur - t0]; }
            sCur = cur;                                                 // (saturating powers: the exact first orbit position beyond the tile)
        }
        __syncthreads();
        for (int L = (int)ORBIT_LEVELS - 2; L >= 0; L--) {
            const uint32_t *src = lv + L * ORBIT_TILE;
            for (uint32_t j = threadIdx.x; j < n; j += 1024)
                if (sF[j]) { const uint32_t y = src[j]; if (y < end) sF[y - t0] = 1; }      // (a mark seen early only marks further orbit positions)
            __syncthreads();
        }
        for (uint32_t j = threadIdx.x; j < n; j += 1024) a.isStart[of + t0 + j] = sF[j];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.isStart[of + a.S] = 0;
}

// bin number of every candidate within its contig (the keys of the dict __binCandidateKmers returns)
__global__ void __launch_bounds__(256) k_pp_bin_numbers(const __grid_constant__ Args a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (p >= a.S) return;
    const size_t o = (size_t)g * a.S, of = (size_t)g * (a.S + 1);
    // binIdx is the exclusive scan: bins that start before p; the bin of p has started at or before p
    const uint32_t mine = a.binIdx[of + p] + a.isStart[of + p] - 1u;
    const uint32_t cf = a.contigFirst[o + p];
    a.binOf[o + a.gid[o + p]] = mine - a.binIdx[of + cf];        // (the contig's first position starts its bin 0)
    if (g == 0 && a.isStart[of + p]) a.binPos[mine] = p;
    if (g == 0 && p == 0) a.binPos[a.binIdx[a.S]] = a.S;          // binIdx[S] = number of bins of the first genome
}

// p1.seq in p2.seq (:34) on 2-bit words
__device__ __forceinline__ bool contains(uint64_t big, int kb, uint64_t small, int ks) {
    const uint64_t mask = ks == 32 ? ~0ull : ((1ull << (2 * ks)) - 1ull);
    for (int sh = 2 * (kb - ks); sh >= 0; sh -= 2)
        if (((big >> sh) & mask) == small) return true;
    return false;
}

// __reduceBinSize (:10-36), one thread per bin of the first genome.  lengths[big] is complete whenever it serves as the
// longer list (a list only loses members while it serves as the shorter one, and those combinations come later in
// combinations(sorted(keys), 2)); removing the current element of the list being iterated makes Python's iterator
// skip the following one: a flag.
__global__ void __launch_bounds__(128) k_pp_reduce(const __grid_constant__ Args a) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nb) return;
    const uint32_t p0 = a.binPos[b], p1 = a.binPos[b + 1];
    uint32_t present = 0;          // bit L - 1: a member of length L
    for (uint32_t p = p0; p < p1; p++) present |= 1u << (a.k[a.gid[p]] - 1);
    for (int small = 1; small <= 32; small++) {
        if (!((present >> (small - 1)) & 1u)) continue;
        for (int big = small + 1; big <= 32; big++) {
            if (!((present >> (big - 1)) & 1u)) continue;
            bool skip = false;
            for (uint32_t p = p0; p < p1; p++) {
                const uint32_t id = a.gid[p];
                if (a.k[id] != small || !a.alive[p]) continue;
                if (skip) { skip = false; continue; }
                const uint64_t w1 = a.word[id];
                for (uint32_t q = p0; q < p1; q++) {
                    const uint32_t id2 = a.gid[q];
                    if (a.k[id2] != big) continue;
                    if (contains(a.word[id2], big, w1, small)) { a.alive[p] = 0; skip = true; break; }
                }
            }
        }
    }
    uint32_t first = 0xFFFFFFFFu, last = 0;
    for (uint32_t p = p0; p < p1; p++)
        if (a.alive[p]) { if (first == 0xFFFFFFFFu) first = p; last = p; }
    a.binFirst[b] = first; a.binLast[b] = last;
}

// __getBinPairs (:101-143) + the generator of __getCandidatePrimerPairs (:283-341): one WARP per first bin b1, one lane per
// second bin b2 > b1 (32 at a time).  The reference leaves the inner loop at the first b2 whose smallest possible product
// is too large (:134; that bound only grows with b2) or at the end of the contig: the lanes from the first such b2 on
// are dropped.  A ballot of the lanes that found a pair gives every pair its place in the reference's order.
template <bool EMIT>
__global__ void __launch_bounds__(256) k_pp_pairs(const __grid_constant__ Args a) {
    const uint32_t b1 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (b1 >= a.nb) return;                                                // (warp-uniform)
    const uint32_t p0 = a.binPos[b1], p1 = a.binPos[b1 + 1];
    const uint32_t c = a.gcontig[p0];
    const uint32_t f0 = a.binFirst[b1], fl = a.binLast[b1];
    const int64_t b1_first_start = a.gleft[f0];
    const int64_t b1_last_end = (int64_t)a.gleft[fl] + a.k[a.gid[fl]] - 1;
    uint32_t n = 0;
    uint32_t out = EMIT ? a.pairOff[b1] : 0u;
    for (uint32_t base = b1 + 1; base < a.nb; base += 32) {
        const uint32_t b2 = base + lane;
        bool stop = b2 >= a.nb, found = false;
        uint32_t pf = 0, pr = 0, ppcr = 0;
        if (!stop) {
            const uint32_t q0 = a.binPos[b2], q1 = a.binPos[b2 + 1];
            if (a.gcontig[q0] != c) stop = true;
            else {
                const uint32_t s0 = a.binFirst[b2], sl = a.binLast[b2];
                const int64_t smallest = ((int64_t)a.gleft[s0] + a.minPrimerLen) - (b1_last_end - a.minPrimerLen);     // :128
                const int64_t largest = ((int64_t)a.gleft[sl] + a.k[a.gid[sl]] - 1) - b1_first_start;                  // :131
                if (smallest > a.maxPcr) stop = true;                                                                  // :134
                else if (largest >= a.minPcr) {                                                                        // :138
                    for (uint32_t p = p0; p < p1 && !found; p++) {
                        if (!a.alive[p]) continue;
                        const uint32_t f = a.gid[p];
                        if (!(a.word[f] & 2ull)) continue;                             // fwd.seq[-1] in {G, C} (:271-272)
                        const int64_t fs = a.gleft[p];
                        const double ftm = a.tm[f];
                        for (uint32_t q = q0; q < q1; q++) {
                            if (!a.alive[q]) continue;
                            const uint32_t r = a.gid[q];
                            const int kr = a.k[r];
                            if (!((a.word[r] >> (2 * (kr - 1))) & 2ull)) continue;     // rev.seq[0] in {G, C} (:273-274)
                            const int64_t pcr = ((int64_t)a.gleft[q] + kr - 1) - fs + 1;                               // :221
                            if (pcr < a.minPcr || pcr > a.maxPcr) continue;                                            // :224
                            if (!(fabs(ftm - a.tm[r]) <= a.maxTmDiff)) continue;                                       // :227
                            pf = f; pr = r; ppcr = (uint32_t)pcr;
                            found = true;
                            break;
                        }
                    }
                }
            }
        }
        const uint32_t stopMask = __ballot_sync(0xFFFFFFFFu, stop);
        const uint32_t valid = stopMask ? ((1u << (__ffs(stopMask) - 1)) - 1u) : 0xFFFFFFFFu;     // the lanes before the first stop
        const uint32_t fmask = __ballot_sync(0xFFFFFFFFu, found) & valid;
        if (EMIT && ((fmask >> lane) & 1u)) {
            const uint32_t at = out + __popc(fmask & ((1u << lane) - 1u));
            a.pairsA[at] = make_uint4(pf, pr, c, ppcr);
            a.pairsB[at] = make_uint2(a.binOf[pf], a.binOf[pr]);
        }
        n += __popc(fmask);
        out += __popc(fmask);
        if (stopMask) break;
    }
    if (!EMIT && lane == 0) a.pairCount[b1] = n;
}

// __getAllSharedPrimerPairs (:367-464), one thread per (pair, genome): the product length, or "no Product here".  The bin
// numbers of :467-502 are binOf[genome][fwd] / binOf[genome][rev]: columns the host already has, not repeated per pair.
constexpr uint32_t PP_VALID = 0x80000000u;
__global__ void __launch_bounds__(256) k_pp_shared(const __grid_constant__ Args a, uint32_t nPairs) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (q >= nPairs) return;
    const uint4 pr = a.pairsA[q];
    uint32_t o = 0u;
    if (g == 0) {
        o = PP_VALID | pr.w;                                                   // out[(p1, p2)][firstName] = product (:398)
    } else {
        const size_t base = (size_t)g * a.S;
        const uint32_t x = pr.x, y = pr.y;
        const int kx = a.k[x], ky = a.k[y];
        // p1.seq is found as it is (k1_rev False); p2.seq = rc of the second candidate's string is found as it is only
        // when that string is its own reverse complement (:403-431)
        uint64_t w = a.word[y], rc = 0;
        for (int i = 0; i < ky; i++) { rc = (rc << 2) | ((w & 3ull) ^ 1ull); w >>= 2; }
        const bool k2_rev = rc != a.word[y];
        const uint32_t sx = a.stOf[base + x], sy = a.stOf[base + y];
        const int64_t lx = a.lfOf[base + x], ly = a.lfOf[base + y];
        if (a.ctOf[base + x] == a.ctOf[base + y] && k2_rev) {                  // :434, :436
            const int64_t s1 = sx ? lx + kx - 1 : lx, s2 = sy ? ly + ky - 1 : ly;     // Primer.start (Primer.py:42-45)
            const bool xfwd = s1 < s2;                                         // :438-445
            if (sx == sy) {                                                    // :448
                const int64_t lf = xfwd ? lx : ly, lr = xfwd ? ly : lx;
                const int kr = xfwd ? ky : kx;
                const int64_t length = (lr + kr - 1) - lf + 1;                 // :455 (both on (+) after :450-452)
                if (length >= a.minPcr && length <= a.maxPcr) o = PP_VALID | (uint32_t)length;      // :456, :463
            }
        }
    }
    a.shared[(size_t)g * nPairs + q] = o;             // genome-major: the threads of a warp write neighbouring words
}

// pairs with a Product in every genome (:459-462)
__global__ void __launch_bounds__(256) k_pp_kept(const __grid_constant__ Args a, uint32_t nPairs) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nPairs) return;
    uint32_t all = PP_VALID;
    for (uint32_t g = 0; g < a.G; g++) all &= a.shared[(size_t)g * nPairs + q];
    a.kept[q] = all ? 1 : 0;
    const uint32_t m = __ballot_sync(__activemask(), all != 0u);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(__activemask()) - 1)) atomicAdd(&a.cnt[4], (uint32_t)__popc(m));
}

// reduced[candidate] for the caller: the alive flags are kept by list position
__global__ void __launch_bounds__(256) k_pp_reduced_by_id(const __grid_constant__ Args a, uint8_t *out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.S) out[a.gid[p]] = a.alive[p] ? 0 : 1;
}

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace pp

struct pfb_pp {
    int device = 0;
    cudaStream_t stream = nullptr, copyStream = nullptr;     // results leave on the second stream while the next kernels run
    cudaEvent_t ev[4] = {}, evPairs = nullptr;
    std::string err;
    uint64_t S = 0;
    uint32_t G = 0, capG = 0;
    pp::Args a{};
    pp::Buf dWord, dK, dTm, dGid, dGcontig, dGleft, dGstrand, ctOf, lfOf, binOf, stOf, seen, jump, isStart, binIdx, contigFirst,
        binPos, alive, binFirst, binLast, pairCount, pairOff, pairsA, pairsB, shared, kept, cnt, tiles, scratch;
    bool ran = false;
    uint64_t nPairs = 0, nKept = 0, nBins = 0, nLaunches = 0;
    float msBin = 0, msPairs = 0, msShared = 0, msTotal = 0;
    // results in pinned host memory
    void *hPairsA = nullptr, *hPairsB = nullptr, *hShared = nullptr, *hKept = nullptr;
    size_t hCapPairs = 0, hCapShared = 0;
};

static std::string g_pp_create_err;

namespace pp {

int fail(pfb_pp *o, int code, const std::string &msg) {
    if (o) o->err = msg; else g_pp_create_err = msg;
    return code;
}

#define PPCK(call)                                                                                     \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return pp::fail(o, PFB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int ensure(pfb_pp *o, Buf &b, size_t bytes) {
    if (b.p && bytes <= b.cap) return PFB_OK;
    if (b.p) PPCK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    PPCK(cudaMalloc(&b.p, want));
    b.cap = want;
    return PFB_OK;
}

// per-genome device arrays with room for capG genomes (grown by doubling; the genomes already copied are kept)
int ensure_genomes(pfb_pp *o, uint32_t want) {
    if (want <= o->capG) return PFB_OK;
    const uint32_t cap = std::max(want, std::max(4u, 2u * o->capG));
    const uint64_t S = o->S;
    Buf *bufs[4] = {&o->dGid, &o->dGcontig, &o->dGleft, &o->dGstrand};
    const size_t elt[4] = {4, 4, 4, 1};
    for (int j = 0; j < 4; j++) {
        Buf nb;
        int rc = ensure(o, nb, (size_t)cap * S * elt[j] + 16);
        if (rc != PFB_OK) return rc;
        if (bufs[j]->p && o->G) PPCK(cudaMemcpy(nb.p, bufs[j]->p, (size_t)o->G * S * elt[j], cudaMemcpyDeviceToDevice));
        if (bufs[j]->p) PPCK(cudaFree(bufs[j]->p));
        *bufs[j] = nb;
    }
    o->capG = cap;
    return PFB_OK;
}

int ensure_pinned(pfb_pp *o, uint64_t nPairs, uint32_t G) {
    if (nPairs > o->hCapPairs) {
        if (o->hPairsA) { cudaFreeHost(o->hPairsA); cudaFreeHost(o->hPairsB); cudaFreeHost(o->hKept); }
        o->hPairsA = o->hPairsB = o->hKept = nullptr; o->hCapPairs = 0;
        const size_t cap = nPairs + nPairs / 4 + 64;
        PPCK(cudaHostAlloc(&o->hPairsA, cap * 16, cudaHostAllocDefault));
        PPCK(cudaHostAlloc(&o->hPairsB, cap * 8, cudaHostAllocDefault));
        PPCK(cudaHostAlloc(&o->hKept, cap, cudaHostAllocDefault));
        o->hCapPairs = cap;
    }
    if (nPairs * G > o->hCapShared) {
        if (o->hShared) cudaFreeHost(o->hShared);
        o->hShared = nullptr; o->hCapShared = 0;
        const size_t cap = nPairs * G + nPairs * G / 4 + 64;
        PPCK(cudaHostAlloc(&o->hShared, cap * 4, cudaHostAllocDefault));
        o->hCapShared = cap;
    }
    return PFB_OK;
}

int run(pfb_pp *o, const pfb_pp_params &prm) {
    const uint64_t S = o->S;
    const uint32_t G = o->G;
    Args &a = o->a;
    int rc;
#define ENS(buf, bytes) if ((rc = ensure(o, o->buf, (bytes))) != PFB_OK) return rc
    ENS(ctOf, (size_t)G * S * 4 + 4); ENS(lfOf, (size_t)G * S * 4 + 4); ENS(binOf, (size_t)G * S * 4 + 4); ENS(stOf, (size_t)G * S + 4);
    ENS(seen, (size_t)G * S * 4 + 4); ENS(jump, (size_t)G * S * 4 + 4); ENS(contigFirst, (size_t)G * S * 4 + 4);
    ENS(isStart, (size_t)G * (S + 1) * 4 + 4); ENS(binIdx, (size_t)G * (S + 1) * 4 + 4);
    ENS(binPos, (S + 2) * 4); ENS(alive, S + 4); ENS(binFirst, (S + 2) * 4); ENS(binLast, (S + 2) * 4);
    ENS(pairCount, (S + 2) * 4); ENS(pairOff, (S + 2) * 4); ENS(cnt, 64); ENS(tiles, (devscan::n_tiles((uint32_t)S + 1) + 4) * 4);
#undef ENS
    a.S = (uint32_t)S; a.G = G;
    a.word = (const uint64_t *)o->dWord.p; a.k = (const uint8_t *)o->dK.p; a.tm = (const double *)o->dTm.p;
    a.gid = (const uint32_t *)o->dGid.p; a.gcontig = (const uint32_t *)o->dGcontig.p; a.gleft = (const uint32_t *)o->dGleft.p;
    a.gstrand = (const uint8_t *)o->dGstrand.p;
    a.ctOf = (uint32_t *)o->ctOf.p; a.lfOf = (uint32_t *)o->lfOf.p; a.binOf = (uint32_t *)o->binOf.p; a.stOf = (uint8_t *)o->stOf.p;
    a.seen = (uint32_t *)o->seen.p; a.jump = (uint32_t *)o->jump.p; a.isStart = (uint32_t *)o->isStart.p; a.binIdx = (uint32_t *)o->binIdx.p;
    a.contigFirst = (uint32_t *)o->contigFirst.p; a.binPos = (uint32_t *)o->binPos.p; a.alive = (uint8_t *)o->alive.p;
    a.binFirst = (uint32_t *)o->binFirst.p; a.binLast = (uint32_t *)o->binLast.p; a.pairCount = (uint32_t *)o->pairCount.p;
    a.pairOff = (uint32_t *)o->pairOff.p; a.cnt = (uint32_t *)o->cnt.p;
    a.maxBin = prm.max_bin; a.minPrimerLen = prm.min_primer_len; a.minPcr = prm.min_pcr; a.maxPcr = prm.max_pcr; a.maxTmDiff = prm.max_tm_diff;
    o->nLaunches = 0; o->nPairs = 0; o->nKept = 0; o->nBins = 0;
    PPCK(cudaEventRecord(o->ev[0], o->stream));
    PPCK(cudaMemsetAsync(o->cnt.p, 0, 64, o->stream));
    if (S == 0 || G == 0) { PPCK(cudaStreamSynchronize(o->stream)); o->ran = true; return PFB_OK; }
    PPCK(cudaMemsetAsync(o->seen.p, 0, (size_t)G * S * 4, o->stream));
    PPCK(cudaMemsetAsync(o->alive.p, 1, S, o->stream));
    const dim3 gridSG((uint32_t)((S + 255) / 256), G);
    k_pp_scatter<<<gridSG, 256, 0, o->stream>>>(a);
    k_pp_contig_first<<<gridSG, 256, 0, o->stream>>>(a);
    k_pp_jump<<<gridSG, 256, 0, o->stream>>>(a);
    k_pp_orbit<<<G, 1024, ORBIT_SMEM, o->stream>>>(a);
    o->nLaunches += 4;
    for (uint32_t g = 0; g < G; g++) {
        devscan::exscan(a.isStart + (size_t)g * (S + 1), (uint32_t)S, a.binIdx + (size_t)g * (S + 1), (uint32_t *)o->tiles.p,
                        a.cnt + (g == 0 ? 0 : 3), o->stream);
        o->nLaunches += 3;
    }
    k_pp_bin_numbers<<<gridSG, 256, 0, o->stream>>>(a);
    o->nLaunches++;
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PPCK(cudaMemcpyAsync(cnt, o->cnt.p, 32, cudaMemcpyDeviceToHost, o->stream));
    PPCK(cudaStreamSynchronize(o->stream));
    if (cnt[2] & ERR_DUP) return fail(o, PFB_EINVAL, "every genome must list every candidate exactly once");
    a.nb = cnt[0]; o->nBins = cnt[0];
    PPCK(cudaEventRecord(o->ev[1], o->stream));
    // first genome: reduction, bin pairs, the first suitable primer pair of each
    k_pp_reduce<<<(a.nb + 127) / 128, 128, 0, o->stream>>>(a);
    k_pp_pairs<false><<<(a.nb + 7) / 8, 256, 0, o->stream>>>(a);
    devscan::exscan(a.pairCount, a.nb, a.pairOff, (uint32_t *)o->tiles.p, a.cnt + 1, o->stream);
    o->nLaunches += 5;
    PPCK(cudaMemcpyAsync(cnt, o->cnt.p, 32, cudaMemcpyDeviceToHost, o->stream));
    PPCK(cudaStreamSynchronize(o->stream));
    const uint64_t nPairs = cnt[1];
    if ((rc = ensure(o, o->pairsA, nPairs * 16 + 16)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->pairsB, nPairs * 8 + 8)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->shared, nPairs * G * 4 + 16)) != PFB_OK) return rc;
    if ((rc = ensure(o, o->kept, nPairs + 16)) != PFB_OK) return rc;
    if ((rc = ensure_pinned(o, nPairs, G)) != PFB_OK) return rc;
    a.pairsA = (uint4 *)o->pairsA.p; a.pairsB = (uint2 *)o->pairsB.p; a.shared = (uint32_t *)o->shared.p; a.kept = (uint8_t *)o->kept.p;
    k_pp_pairs<true><<<(a.nb + 7) / 8, 256, 0, o->stream>>>(a);
    o->nLaunches++;
    PPCK(cudaEventRecord(o->ev[2], o->stream));
    if (nPairs) {
        PPCK(cudaEventRecord(o->evPairs, o->stream));
        PPCK(cudaStreamWaitEvent(o->copyStream, o->evPairs, 0));
        PPCK(cudaMemcpyAsync(o->hPairsA, o->pairsA.p, nPairs * 16, cudaMemcpyDeviceToHost, o->copyStream));     // (beside the look-up below)
        PPCK(cudaMemcpyAsync(o->hPairsB, o->pairsB.p, nPairs * 8, cudaMemcpyDeviceToHost, o->copyStream));
        k_pp_shared<<<dim3((uint32_t)((nPairs + 255) / 256), G), 256, 0, o->stream>>>(a, (uint32_t)nPairs);
        k_pp_kept<<<(uint32_t)((nPairs + 255) / 256), 256, 0, o->stream>>>(a, (uint32_t)nPairs);
        o->nLaunches += 2;
    }
    PPCK(cudaEventRecord(o->ev[3], o->stream));
    if (nPairs) {
        PPCK(cudaMemcpyAsync(o->hShared, o->shared.p, nPairs * G * 4, cudaMemcpyDeviceToHost, o->stream));
        PPCK(cudaMemcpyAsync(o->hKept, o->kept.p, nPairs, cudaMemcpyDeviceToHost, o->stream));
    }
    PPCK(cudaMemcpyAsync(cnt, o->cnt.p, 32, cudaMemcpyDeviceToHost, o->stream));
    PPCK(cudaStreamSynchronize(o->stream));
    PPCK(cudaStreamSynchronize(o->copyStream));
    PPCK(cudaGetLastError());
    o->nPairs = nPairs; o->nKept = cnt[4];
    PPCK(cudaEventElapsedTime(&o->msBin, o->ev[0], o->ev[1]));
    PPCK(cudaEventElapsedTime(&o->msPairs, o->ev[1], o->ev[2]));
    PPCK(cudaEventElapsedTime(&o->msShared, o->ev[2], o->ev[3]));
    o->msTotal = o->msBin + o->msPairs + o->msShared;
    o->ran = true;
    return PFB_OK;
}

}  // namespace pp

extern "C" {

int pfb_pp_create(pfb_pp **out, int device) {
    if (!out) return PFB_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return pp::fail(nullptr, PFB_ECUDA, "no CUDA device (primerforge_b200 has no CPU path)");
    if (device < 0 || device >= n) return pp::fail(nullptr, PFB_EINVAL, "no such device");
    pfb_pp *ctx = new pfb_pp();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evPairs, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return pp::fail(nullptr, PFB_ECUDA, "cudaSetDevice / cudaStreamCreate failed");
    }
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return pp::fail(nullptr, PFB_ECUDA, "cudaEventCreate failed"); }
    if (cudaFuncSetAttribute(pp::k_pp_orbit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp::ORBIT_SMEM) != cudaSuccess) {
        delete ctx;
        return pp::fail(nullptr, PFB_ECUDA, "this device cannot give a CTA 100 KB of shared memory");
    }
    *out = ctx;
    return PFB_OK;
}

void pfb_pp_destroy(pfb_pp *o) {
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->stream) cudaStreamSynchronize(o->stream);
    for (pp::Buf *b : {&o->dWord, &o->dK, &o->dTm, &o->dGid, &o->dGcontig, &o->dGleft, &o->dGstrand, &o->ctOf, &o->lfOf, &o->binOf,
                       &o->stOf, &o->seen, &o->jump, &o->isStart, &o->binIdx, &o->contigFirst, &o->binPos, &o->alive, &o->binFirst,
                       &o->binLast, &o->pairCount, &o->pairOff, &o->pairsA, &o->pairsB, &o->shared, &o->kept, &o->cnt, &o->tiles, &o->scratch})
        if (b->p) cudaFree(b->p);
    for (void *h : {o->hPairsA, o->hPairsB, o->hShared, o->hKept}) if (h) cudaFreeHost(h);
    for (auto &e : o->ev) if (e) cudaEventDestroy(e);
    if (o->evPairs) cudaEventDestroy(o->evPairs);
    if (o->copyStream) cudaStreamDestroy(o->copyStream);
    if (o->stream) cudaStreamDestroy(o->stream);
    delete o;
}

const char *pfb_pp_last_error(const pfb_pp *o) { return o ? o->err.c_str() : g_pp_create_err.c_str(); }

int pfb_pp_set_candidates(pfb_pp *o, const uint64_t *word, const uint8_t *k, const double *tm, uint64_t n) {
    if (!o) return PFB_EINVAL;
    if (n && (!word || !k || !tm)) return pp::fail(o, PFB_EINVAL, "bad candidate arguments");
    if (n >= (1ull << 31)) return pp::fail(o, PFB_ETOOLARGE, "too many candidates");
    for (uint64_t i = 0; i < n; i++)
        if (k[i] < 1 || k[i] > 32) return pp::fail(o, PFB_EINVAL, "candidate length must be 1..32");
    if (cudaSetDevice(o->device) != cudaSuccess) return pp::fail(o, PFB_ECUDA, "cudaSetDevice failed");
    int rc;
    if ((rc = pp::ensure(o, o->dWord, n * 8 + 8)) != PFB_OK || (rc = pp::ensure(o, o->dK, n + 8)) != PFB_OK ||
        (rc = pp::ensure(o, o->dTm, n * 8 + 8)) != PFB_OK) return rc;
    if (n) {      // the columns go to the device as they are: the caller's buffers are free again when the call returns
        PPCK(cudaMemcpy(o->dWord.p, word, n * 8, cudaMemcpyHostToDevice));
        PPCK(cudaMemcpy(o->dK.p, k, n, cudaMemcpyHostToDevice));
        PPCK(cudaMemcpy(o->dTm.p, tm, n * 8, cudaMemcpyHostToDevice));
    }
    o->S = n; o->G = 0; o->capG = 0;
    for (pp::Buf *b : {&o->dGid, &o->dGcontig, &o->dGleft, &o->dGstrand}) { if (b->p) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    o->ran = false;
    return PFB_OK;
}

int pfb_pp_add_genome(pfb_pp *o, const uint32_t *id, const uint32_t *contig, const uint32_t *left, const uint8_t *strand, uint64_t n) {
    if (!o) return PFB_EINVAL;
    if (n != o->S) return pp::fail(o, PFB_EINVAL, "every genome lists every candidate exactly once (getCandidateKmers.py:442-489)");
    if (n && (!id || !contig || !left || !strand)) return pp::fail(o, PFB_EINVAL, "bad genome arguments");
    for (uint64_t p = 1; p < n; p++)
        if (contig[p] < contig[p - 1]) return pp::fail(o, PFB_EINVAL, "contig numbers must follow the order of the lists");
    if (o->G == 0)
        for (uint64_t p = 0; p < n; p++)
            if (strand[p]) return pp::fail(o, PFB_EINVAL, "the first genome's candidates are all on (+) (getCandidateKmers.py:58-63)");
    if (cudaSetDevice(o->device) != cudaSuccess) return pp::fail(o, PFB_ECUDA, "cudaSetDevice failed");
    int rc = pp::ensure_genomes(o, o->G + 1);
    if (rc != PFB_OK) return rc;
    const size_t at = (size_t)o->G * n;
    if (n) {
        PPCK(cudaMemcpy((uint32_t *)o->dGid.p + at, id, n * 4, cudaMemcpyHostToDevice));
        PPCK(cudaMemcpy((uint32_t *)o->dGcontig.p + at, contig, n * 4, cudaMemcpyHostToDevice));
        PPCK(cudaMemcpy((uint32_t *)o->dGleft.p + at, left, n * 4, cudaMemcpyHostToDevice));
        PPCK(cudaMemcpy((uint8_t *)o->dGstrand.p + at, strand, n, cudaMemcpyHostToDevice));
    }
    o->G++;
    o->ran = false;
    return PFB_OK;
}

int pfb_pp_clear_genomes(pfb_pp *o) {
    if (!o) return PFB_EINVAL;
    o->G = 0;
    o->ran = false;
    return PFB_OK;
}

int pfb_pp_run(pfb_pp *o, const pfb_pp_params *prm) {
    if (!o || !prm) return PFB_EINVAL;
    if (prm->max_bin < 0 || prm->min_pcr > prm->max_pcr) return pp::fail(o, PFB_EINVAL, "bad pair-search parameters");
    if (cudaSetDevice(o->device) != cudaSuccess) return pp::fail(o, PFB_ECUDA, "cudaSetDevice failed");
    return pp::run(o, *prm);
}

int pfb_pp_counts(pfb_pp *o, uint64_t *n_pairs, uint64_t *n_kept, uint64_t *n_bins) {
    if (!o) return PFB_EINVAL;
    if (!o->ran) return pp::fail(o, PFB_ESTATE, "no run yet");
    if (n_pairs) *n_pairs = o->nPairs;
    if (n_kept) *n_kept = o->nKept;
    if (n_bins) *n_bins = o->nBins;
    return PFB_OK;
}

int pfb_pp_pairs(pfb_pp *o, pfb_pp_pair *out, uint64_t cap) {
    if (!o || (!out && cap)) return PFB_EINVAL;
    if (!o->ran) return pp::fail(o, PFB_ESTATE, "no run yet");
    if (cap < o->nPairs) return pp::fail(o, PFB_EINVAL, "pair buffer too small");
    const uint4 *A = (const uint4 *)o->hPairsA;
    const uint2 *B = (const uint2 *)o->hPairsB;
    for (uint64_t q = 0; q < o->nPairs; q++) {
        out[q].fwd = A[q].x; out[q].rev = A[q].y; out[q].contig = A[q].z; out[q].fbin = B[q].x; out[q].rbin = B[q].y; out[q].pcr_len = A[q].w;
    }
    return PFB_OK;
}

int pfb_pp_shared(pfb_pp *o, uint32_t *out, uint64_t cap) {
    if (!o || (!out && cap)) return PFB_EINVAL;
    if (!o->ran) return pp::fail(o, PFB_ESTATE, "no run yet");
    const uint64_t n = o->nPairs * o->G;
    if (cap < n) return pp::fail(o, PFB_EINVAL, "buffer too small");
    if (n) memcpy(out, o->hShared, n * 4);
    return PFB_OK;
}

int pfb_pp_kept(pfb_pp *o, uint8_t *out, uint64_t cap) {
    if (!o || (!out && cap)) return PFB_EINVAL;
    if (!o->ran) return pp::fail(o, PFB_ESTATE, "no run yet");
    if (cap < o->nPairs) return pp::fail(o, PFB_EINVAL, "buffer too small");
    if (o->nPairs) memcpy(out, o->hKept, o->nPairs);
    return PFB_OK;
}

int pfb_pp_bins(pfb_pp *o, uint32_t genome, uint32_t *bin_of, uint8_t *reduced, uint64_t cap) {
    if (!o || !bin_of) return PFB_EINVAL;
    if (!o->ran) return pp::fail(o, PFB_ESTATE, "no run yet");
    if (genome >= o->G || cap < o->S) return pp::fail(o, PFB_EINVAL, "bad genome / buffer too small");
    if (cudaSetDevice(o->device) != cudaSuccess) return pp::fail(o, PFB_ECUDA, "cudaSetDevice failed");
    const uint64_t S = o->S;
    if (!S) return PFB_OK;
    PPCK(cudaMemcpy(bin_of, (uint32_t *)o->binOf.p + (size_t)genome * S, S * 4, cudaMemcpyDeviceToHost));
    if (reduced) {
        if (genome != 0) return pp::fail(o, PFB_EINVAL, "only the first genome's bins are reduced");
        int rc = pp::ensure(o, o->scratch, S + 16);
        if (rc != PFB_OK) return rc;
        pp::k_pp_reduced_by_id<<<(uint32_t)((S + 255) / 256), 256, 0, o->stream>>>(o->a, (uint8_t *)o->scratch.p);
        PPCK(cudaStreamSynchronize(o->stream));
        PPCK(cudaMemcpy(reduced, o->scratch.p, S, cudaMemcpyDeviceToHost));
    }
    return PFB_OK;
}

int pfb_pp_timings(pfb_pp *o, pfb_pp_timing *out) {
    if (!o || !out) return PFB_EINVAL;
    memset(out, 0, sizeof(*out));
    out->ms_bin = o->msBin; out->ms_pairs = o->msPairs; out->ms_shared = o->msShared; out->ms_total = o->msTotal;
    out->n_candidates = o->S; out->n_genomes = o->G; out->n_bins = o->nBins; out->n_pairs = o->nPairs; out->n_kept = o->nKept;
    out->n_launches = o->nLaunches;
    return PFB_OK;
}

}  // extern "C"
