// pfb_fasta.cuh -- FASTA file bytes -> contigs, on the device (included by pfb200.cu).
//
// Replaces Bio.SeqIO.parse(fn, "fasta") at /root/reference/bin/getCandidateKmers.py:93,212 (biopython's
// SimpleFastaParser): text before the first line that starts with '>' is skipped; a line that starts with
// '>' opens a record (its id is the title up to the first whitespace -- extracted by the host from the
// header offset reported here); every other line is rstrip()ped -- trailing white space of any kind goes -- and
// contributes what is left minus ' ' and '\r' (SimpleFastaParser: lines.append(line.rstrip()), then
// "".join(lines).replace(" ", "").replace("\r", "")); a tab INSIDE a line stays a byte of the sequence.  Case is
// preserved (only upper-case A C G T are nucleotides for the stage, like on the pfb_add_genome path).
//
// The files of a batch sit in one staging buffer (each 256-byte aligned).  A TILE of 4096 bytes is walked
// by one warp, 32 bytes per step: the line-start / header / newline masks of a step are warp ballots, i.e.
// uniform values, and a lane owns one byte, so the compacted bases leave as coalesced byte stores.
//   k_fa_mark   : per tile: position of its last line start, number of header lines, first header of the file
//   k_fa_scan   : exclusive MAX scan of the last line starts (which line does a tile begin inside?) and
//                 exclusive SUM scan of the header counts / of the kept-byte counts over the tiles
//   k_fa_sweep<false> : kept bytes per tile;  k_fa_sweep<true> : compaction + one entry per record
// Traffic: the file bytes are read three times (they fit L2 for bacterial genomes) and written once.
//
// GenBank flat files (the reference's default --format; Bio.SeqIO.parse(fn, "genbank")) go through the same
// machinery with other line rules: a line starting with "LOCUS" opens a record (the host takes the id from its
// VERSION / LOCUS lines), the lines between an "ORIGIN" line and the next "//" (or LOCUS) line are the sequence,
// of which only letters are kept, upper-cased.  A byte's fate then depends on the last MARK line (ORIGIN, //,
// LOCUS) before it and on whether a line started after that mark: two exclusive MAX scans over the tiles.
#pragma once

#include <cstdint>

namespace fa {

constexpr uint32_t TILE = 4096;        // bytes per warp
constexpr uint32_t WARPS = 8;          // warps (= tiles) per CTA
constexpr uint32_t NO_HDR = 0xFFFFFFFFu;

enum { FMT_FASTA = 1, FMT_GENBANK = 2 };

struct Args {
    int format;                 // FMT_FASTA / FMT_GENBANK
    const uint8_t *stage;       // staging buffer holding the files of the batch
    const uint64_t *src;        // [nFiles] offset of the file in `stage`
    const uint64_t *len;        // [nFiles] bytes
    const uint32_t *tile0;      // [nFiles + 1] first tile of the file
    const uint64_t *dst;        // [nFiles] offset of the file's bases in `out`
    uint32_t nFiles, nTiles;
    // per tile
    uint64_t *tLast;            // (file << 32 | offset of the last line start + 1), 0 = no line start in the tile
    uint64_t *inLast;           // exclusive max scan of tLast
    uint64_t *tMark, *inMark;   // GenBank: the same for the last ORIGIN / "//" / LOCUS line
    uint32_t *tHdr, *hdrBase;   // header lines in the tile / before it (all files of the batch), [nTiles + 1]
    uint32_t *tKept, *keptBase; // kept bytes in the tile / before it, [nTiles + 1]
    uint32_t *firstHdr;         // [nFiles] offset of the file's first header line (NO_HDR: none)
    // output
    uint8_t *out;               // compacted bases (the raw arena)
    uint32_t *recStart;         // [cap] offset of the record's first base in the file's compacted bases (host-mapped)
    uint32_t *recHdr;           // [cap] offset of the record's '>' in the file (host-mapped)
    uint32_t *fileTab;          // [2 * nFiles + 1] per file: index of its first record, then kept bytes; last: total records
    uint32_t cap;
};

__device__ __forceinline__ uint32_t file_of_tile(const Args &a, uint32_t t) {
    uint32_t lo = 0, hi = a.nFiles;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (a.tile0[mid] <= t) lo = mid; else hi = mid;
    }
    // files without tiles (length 0) share their tile0 with the next one: take the last file that starts here
    return lo;
}

// masks of one 32-byte step (bit i = byte i of the step); `c` is this lane's byte
struct Step {
    uint32_t valid, nl, drop, ls, hs;   // hs: lines that open a record ('>' / LOCUS)
    uint32_t origin, off;               // GenBank: ORIGIN lines; "//" and LOCUS lines (sequence off)
    uint32_t alpha;                     // GenBank: letters
    uint32_t tws;                       // FASTA: white space other than ' ' '\r' '\n' (dropped only at the end of a line)
};

__device__ __forceinline__ bool starts_with(const uint8_t *f, uint64_t off, uint64_t len, const char *word, int n) {
    if (off + (uint64_t)n > len) return false;
    for (int i = 0; i < n; i++)
        if (f[off + i] != (uint8_t)word[i]) return false;
    return true;
}

template <int FORMAT>
__device__ __forceinline__ Step classify(const uint8_t *f, uint64_t off, uint64_t len, uint8_t c, bool in_file, uint32_t &prev_nl) {
    Step s;
    const uint32_t lane = threadIdx.x & 31;
    s.valid = __ballot_sync(0xFFFFFFFFu, in_file);
    s.nl = __ballot_sync(0xFFFFFFFFu, in_file && c == '\n');
    s.ls = ((s.nl << 1) | prev_nl) & s.valid;      // a line starts after every newline (and at offset 0 of the file)
    prev_nl = s.nl >> 31;
    s.origin = s.off = s.alpha = s.tws = 0;
    if (FORMAT == FMT_FASTA) {
        s.drop = __ballot_sync(0xFFFFFFFFu, in_file && (c == '\r' || c == ' '));
        // str.isspace() bytes besides those: \t \v \f and the four separators 0x1c..0x1f
        s.tws = __ballot_sync(0xFFFFFFFFu, in_file && (c == '\t' || c == 0x0b || c == 0x0c || (uint8_t)(c - 0x1c) < 4));
        s.hs = s.ls & __ballot_sync(0xFFFFFFFFu, in_file && c == '>');
    } else {
        s.drop = 0;
        const bool at_ls = (s.ls >> lane) & 1u;
        // the three kinds of line that matter start with a distinctive first byte: look further only there
        const bool locus = at_ls && c == 'L' && starts_with(f, off, len, "LOCUS", 5);
        const bool origin = at_ls && c == 'O' && starts_with(f, off, len, "ORIGIN", 6);
        const bool slash = at_ls && c == '/' && starts_with(f, off, len, "//", 2);
        s.hs = __ballot_sync(0xFFFFFFFFu, locus);
        s.origin = __ballot_sync(0xFFFFFFFFu, origin);
        s.off = __ballot_sync(0xFFFFFFFFu, slash) | s.hs;
        s.alpha = __ballot_sync(0xFFFFFFFFu, in_file && (uint8_t)((c | 0x20) - 'a') < 26);
    }
    return s;
}

// FASTA: the white-space bytes of a step that line.rstrip() removes: those followed by nothing but white space up
// to the end of their line (or of the file).  Steps without such bytes -- all of them in a regular file -- cost
// one test; a run that reaches the end of the step is resolved by reading ahead (every lane reads the same bytes).
__device__ __forceinline__ uint32_t trailing_ws(const uint8_t *f, uint64_t len, uint64_t sb, const Step &s) {
    if (!s.tws) return 0u;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t W = s.tws | s.drop;                      // white space that is not a newline
    const uint32_t solid = s.valid & ~W & ~s.nl;            // bytes that end a run of white space without ending the line
    const uint32_t above = lane == 31 ? 0u : (0xFFFFFFFEu << lane);
    const uint32_t stop = (solid | s.nl | ~s.valid) & above;   // first byte after this lane that is not white space (or is off the file)
    const bool mine = (s.tws >> lane) & 1u;
    bool trail = false;
    if (mine && stop) trail = !((solid >> (__ffs(stop) - 1)) & 1u);
    if (__ballot_sync(0xFFFFFFFFu, mine && !stop)) {         // a run touches the end of the step: look ahead
        uint64_t x = sb + 32;
        bool ends_line = true;
        for (; x < len; x++) {
            const uint8_t c = f[x];
            if (c == '\n') break;
            if (!(c == ' ' || c == '\r' || c == '\t' || c == 0x0b || c == 0x0c || (uint8_t)(c - 0x1c) < 4)) { ends_line = false; break; }
        }
        if (mine && !stop) trail = ends_line;
    }
    return __ballot_sync(0xFFFFFFFFu, trail);
}

// is the byte before `off` a newline (or off the start of the file)?
__device__ __forceinline__ uint32_t starts_line(const uint8_t *f, uint64_t off) { return off == 0 || f[off - 1] == '\n'; }

template <int FORMAT>
__global__ void __launch_bounds__(WARPS * 32) k_fa_mark(const Args a) {
    const uint32_t t = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= a.nTiles) return;
    const uint32_t fi = file_of_tile(a, t);
    const uint8_t *f = a.stage + a.src[fi];
    const uint64_t len = a.len[fi];
    const uint64_t base = (uint64_t)(t - a.tile0[fi]) * TILE;
    uint32_t prev_nl = starts_line(f, base);
    uint32_t last = 0, mark = 0, nh = 0, first = NO_HDR;
    for (uint32_t j = 0; j < TILE / 32; j++) {
        const uint64_t off = base + j * 32 + lane;
        const bool in_file = off < len;
        const uint8_t c = in_file ? f[off] : 0;
        const Step s = classify<FORMAT>(f, off, len, c, in_file, prev_nl);
        if (s.ls) last = (uint32_t)(base + j * 32) + (31 - __clz(s.ls)) + 1;
        const uint32_t mk = s.origin | s.off;
        if (mk) mark = (uint32_t)(base + j * 32) + (31 - __clz(mk)) + 1;
        if (s.hs) {
            if (first == NO_HDR) first = (uint32_t)(base + j * 32) + (__ffs(s.hs) - 1);
            nh += __popc(s.hs);
        }
    }
    if (lane == 0) {
        a.tLast[t] = last ? (((uint64_t)fi << 32) | last) : 0ull;
        if (FORMAT == FMT_GENBANK) a.tMark[t] = mark ? (((uint64_t)fi << 32) | mark) : 0ull;
        a.tHdr[t] = nh;
        if (first != NO_HDR) atomicMin(&a.firstHdr[fi], first);
    }
}

// single CTA: exclusive scans over the tiles: SUM of val (n + 1 outputs), MAX of mval and of mval2 (each optional)
__global__ void __launch_bounds__(1024) k_fa_scan(const uint32_t *val, uint32_t *out, const uint64_t *mval, uint64_t *mout,
                                                  const uint64_t *mval2, uint64_t *mout2, uint32_t n) {
    __shared__ uint32_t ssum[1024];
    __shared__ uint64_t smax[1024], smax2[1024];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    uint32_t s = 0;
    uint64_t m = 0, m2 = 0;
    for (uint32_t x = lo; x < hi; x++) {
        s += val[x];
        if (mval) m = max(m, mval[x]);
        if (mval2) m2 = max(m2, mval2[x]);
    }
    ssum[threadIdx.x] = s; smax[threadIdx.x] = m; smax2[threadIdx.x] = m2;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= off ? ssum[threadIdx.x - off] : 0;
        uint64_t w = threadIdx.x >= off ? smax[threadIdx.x - off] : 0;
        uint64_t w2 = threadIdx.x >= off ? smax2[threadIdx.x - off] : 0;
        __syncthreads();
        ssum[threadIdx.x] += v; smax[threadIdx.x] = max(smax[threadIdx.x], w); smax2[threadIdx.x] = max(smax2[threadIdx.x], w2);
        __syncthreads();
    }
    uint32_t run = ssum[threadIdx.x] - s;
    uint64_t mrun = threadIdx.x ? smax[threadIdx.x - 1] : 0, mrun2 = threadIdx.x ? smax2[threadIdx.x - 1] : 0;
    for (uint32_t x = lo; x < hi; x++) {
        uint32_t v = val[x];
        out[x] = run; run += v;
        if (mval) { uint64_t w = mval[x]; mout[x] = mrun; mrun = max(mrun, w); }
        if (mval2) { uint64_t w = mval2[x]; mout2[x] = mrun2; mrun2 = max(mrun2, w); }
    }
    if (threadIdx.x == 1023) out[n] = ssum[1023];
}

// GenBank: what the bytes of the CURRENT line are, given the last mark line at or before it
enum { GB_OFF = 0, GB_ORIGIN_LINE = 1, GB_SEQ = 2 };

// GATHER = false: kept bytes per tile.  GATHER = true: compaction and the record table.
template <int FORMAT, bool GATHER>
__global__ void __launch_bounds__(WARPS * 32) k_fa_sweep(const Args a) {
    const uint32_t t = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= a.nTiles) return;
    const uint32_t fi = file_of_tile(a, t);
    const uint8_t *f = a.stage + a.src[fi];
    const uint64_t len = a.len[fi];
    const uint32_t t0 = a.tile0[fi];
    const uint64_t base = (uint64_t)(t - t0) * TILE;
    const uint32_t F = a.firstHdr[fi];
    uint32_t prev_nl = starts_line(f, base);
    // the line the tile begins inside (or, when the tile begins a line, the line before it) started at the last
    // line start before the tile -- same file by construction, offset 0 of a file starts a line.
    // FASTA: is that line a header line?  GenBank: the state the last mark line left behind.
    uint32_t state = 0;
    if (t > t0) {
        const uint32_t il = (uint32_t)a.inLast[t];
        if (FORMAT == FMT_FASTA) {
            state = f[il - 1u] == '>';
        } else {
            const uint64_t mk = a.inMark[t];
            if ((uint32_t)(mk >> 32) == fi && (uint32_t)mk != 0u && f[(uint32_t)mk - 1u] == 'O')
                state = il > (uint32_t)mk ? GB_SEQ : GB_ORIGIN_LINE;
        }
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t run = 0, hrun = 0;
    uint8_t *out = nullptr;
    if (GATHER) {
        run = a.keptBase[t] - a.keptBase[t0];            // offset in the file's compacted bases
        hrun = a.hdrBase[t];
        out = a.out + a.dst[fi];
        if (t == t0 && lane == 0) {
            a.fileTab[2 * fi] = a.hdrBase[t0];
            a.fileTab[2 * fi + 1] = a.keptBase[a.tile0[fi + 1]] - a.keptBase[t0];
            if (fi == 0) a.fileTab[2 * a.nFiles] = a.hdrBase[a.nTiles];
        }
    }
    for (uint32_t j = 0; j < TILE / 32; j++) {
        const uint64_t sb = base + j * 32;
        const uint64_t off = sb + lane;
        const bool in_file = off < len;
        const uint8_t c = in_file ? f[off] : 0;
        const Step s = classify<FORMAT>(f, off, len, c, in_file, prev_nl);
        // a line keeps the status decided at its start: `on` = bytes of lines in the "on" status
        // (FASTA: header lines; GenBank: sequence lines)
        uint32_t on = 0, prev = 0;
        for (uint32_t m = s.ls; m; m &= m - 1) {
            const uint32_t b = __ffs(m) - 1;
            const bool is_on = FORMAT == FMT_FASTA ? state != 0 : state == GB_SEQ;
            if (is_on && b > prev) on |= (0xFFFFFFFFu >> (32 - (b - prev))) << prev;
            if (FORMAT == FMT_FASTA) state = (s.hs >> b) & 1u;
            else state = ((s.origin >> b) & 1u) ? GB_ORIGIN_LINE : ((s.off >> b) & 1u) ? GB_OFF : (state == GB_OFF ? GB_OFF : GB_SEQ);
            prev = b;
        }
        if (FORMAT == FMT_FASTA ? state != 0 : state == GB_SEQ) on |= 0xFFFFFFFFu << prev;
        // bytes before the first record of the file are not part of any record
        uint32_t junk = 0;
        if (F == NO_HDR || sb + 32 <= F) junk = 0xFFFFFFFFu;
        else if (sb < F) junk = 0xFFFFFFFFu >> (32 - (uint32_t)(F - sb));
        uint32_t keep = FORMAT == FMT_FASTA ? (s.valid & ~on & ~s.nl & ~s.drop & ~junk) : (s.valid & on & s.alpha & ~junk);
        if (FORMAT == FMT_FASTA) keep &= ~trailing_ws(f, len, sb, s);
        if (GATHER) {
            const uint32_t pos = run + __popc(keep & lt_mask);
            if ((keep >> lane) & 1u) out[pos] = FORMAT == FMT_FASTA ? c : (uint8_t)(c & 0xDF);   // GenBank sequences are upper-cased
            if ((s.hs >> lane) & 1u) {
                const uint32_t e = hrun + __popc(s.hs & lt_mask);
                if (e < a.cap) { a.recStart[e] = pos; a.recHdr[e] = (uint32_t)off; }
            }
            hrun += __popc(s.hs);
        }
        run += __popc(keep);
    }
    if (!GATHER && lane == 0) a.tKept[t] = run;
}

}  // namespace fa
