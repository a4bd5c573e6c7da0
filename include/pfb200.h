/*
 * pfb200.h -- C ABI of libpfb200.so, the B200 (sm_100a) implementation of primerForge's
 * candidate-k-mer stage.
 *
 * The reference has no FFI: the stage is the Python function
 *   _getAllCandidateKmers(params, sharedExists)        /root/reference/bin/getCandidateKmers.py:517
 * called from bin/main.py:86.  These entry points are what a ctypes binding behind that function
 * binds (see INTEGRATION.md and primerforge_b200/_lib.py); each one cites the reference code it
 * replaces.  Plain pointers and sizes only; the library owns all device memory, the caller owns
 * all host buffers.  Every function returns 0 on success or a negative PFB_E* code and never
 * throws; pfb_last_error() gives the message.  A context is not thread-safe; it lives on one GPU.
 * More than one GPU: either one process per GPU (pfb_comm_*: every rank joins a communicator and calls
 * pfb_run / pfb_compute like a single GPU would; the exchanges happen inside) or one process for all
 * GPUs (pfb_multi_*: one host thread drives N contexts).
 */
#ifndef PFB200_H
#define PFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFB_OK            0
#define PFB_EINVAL       -1   /* bad argument / parameter out of range */
#define PFB_ECUDA        -2   /* CUDA runtime error (no device, out of memory, launch failure) */
#define PFB_ESTATE       -3   /* call sequence error */
#define PFB_ENOSHARED    -4   /* no k-mer is shared by all ingroup genomes (getCandidateKmers.py:566-568) */
#define PFB_ETOOLARGE    -5   /* a genome exceeds 2^26 padded bases */

typedef struct pfb_ctx pfb_ctx;

/* Stage parameters: the fields of bin/Parameters.py the stage reads plus the
 * constants of the third-party code it calls. */
typedef struct pfb_params {
    int32_t  k_min, k_max;        /* params.minLen / maxLen (getCandidateKmers.py:103); 1 <= k_min <= k_max <= 32 */
    uint32_t table_size;          /* khmer Countgraph(k, 1e7, 1) table prime = 9999991 (:46-47) */
    int32_t  max_repeat;          /* params.maxRepeatLen: homopolymer length that is rejected (:406-409) */
    double   min_gc, max_gc;      /* params.minGc / maxGc, inclusive (:303) */
    double   min_tm, max_tm;      /* params.minTm / maxTm, inclusive (:307) */
    /* primer3.calc_tm defaults (bin/Primer.py:104 passes none): mM, mM, mM, nM, %, -, mol/l */
    double   mv_conc, dv_conc, dntp_conc, dna_conc, dmso_conc, dmso_fact, formamide_conc;
    int32_t  prefilter;           /* 1: apply GC / repeat / Tm to genome 1's k-mers BEFORE the cross-genome join
                                     (records are only the k-mers that pass; same candidate set, fewer records);
                                     0: report every shared k-mer with its filter flags (== the reference's
                                     sharedKmers dict, getCandidateKmers.py:190-219) */
    int32_t  reserved;
} pfb_params;

/* One shared k-mer, in the reference's dict insertion order (genome-1 contig, end ascending,
 * longest first; getCandidateKmers.py:171-187 + pyahocorasick's emission order). */
typedef struct pfb_record {
    uint64_t word;      /* the k-mer as it reads on genome 1's (+) strand: 2 bits per base, A0 T1 C2 G3,
                           first base most significant */
    double   tm;        /* primer3.calc_tm equivalent (Primer.py:102-104) */
    uint32_t slot;      /* internal id (stable across the run) */
    uint16_t gc_count;  /* gcPer = gc_count / k * 100 (Primer.py:89-100) */
    uint8_t  k;
    uint8_t  flags;     /* PFB_F_* */
} pfb_record;

#define PFB_F_GC_OK     1u   /* minGc <= gcPer <= maxGc  (getCandidateKmers.py:301-303) */
#define PFB_F_TM_OK     2u   /* minTm <= Tm <= maxTm     (:305-307) */
#define PFB_F_REP_OK    4u   /* no homopolymer >= maxRepeatLen (:309-320) */
#define PFB_F_PICK      8u   /* first k-mer passing the three filters at its (contig, start) in at least one
                                genome (:376-388 with hairpin / homodimer left to the caller) */
#define PFB_F_PALIN    16u   /* reverse-complement palindrome */

/* Position of one record in one genome (getCandidateKmers.py:173-187): contig index in the order
 * the contigs were added, leftmost (+)-strand coordinate, strand (0 '+', 1 '-'). */
typedef struct pfb_pos {
    uint32_t contig;
    uint32_t start;
    uint32_t strand;
} pfb_pos;

/* CUDA-event timings of the last pfb_run (milliseconds) and the counts the roofline uses */
typedef struct pfb_timing {
    float    ms_upload;        /* host -> device copies of the ASCII bases */
    float    ms_encode;        /* 2-bit packing */
    float    ms_first;         /* genome-1 candidate table + key table build */
    float    ms_scan;          /* every genome against the key table (the hot kernel) */
    float    ms_finalize;      /* per-genome uniqueness rules, ordered emit, grouping, positions */
    float    ms_total;         /* encode .. finalize (device resident -> device resident) */
    uint64_t n_bases;          /* sum of ingroup bases on this context */
    uint64_t n_windows;        /* k-mer windows hashed (all genomes, all k) */
    uint64_t n_slots;          /* genome-1 k-mers entered into the key table */
    uint64_t n_records;        /* shared k-mers reported */
    uint64_t n_picked;
    uint64_t n_launches;       /* kernels launched by the last run */
    uint64_t used_filter;      /* 1: the scan probed through the shared-memory filter (k_scan_filtered) */
    uint64_t scan_main_bases;  /* bases covered by the one scan launch over genomes 2..G (pfb_compute only) */
    float    ms_scan_main;     /* CUDA-event duration of that launch: the roofline kernel of bench.py */
    float    ms_exchange;      /* CUDA-event time of this run's collectives on this context's stream (0 on one GPU) */
    uint64_t n_ids;            /* genome-1 k-mers that are unique in genome 1 (= length of the per-id result columns) */
    uint64_t sharded_first_genome; /* 1: the ranks shared genome 1's own pipeline (candidates + its scan) by position */
    uint64_t n_reruns;         /* runs of this context that were repeated because a buffer capacity proved too small */
    uint64_t exchange_mode;    /* 0 one GPU; 1 NCCL collectives; 2 the ranks read each other's exchange arenas over NVLink */
} pfb_timing;

/* The results as the device keeps them: one entry per id.  ids are the genome-1 k-mers that are unique in
 * genome 1, numbered in the reference's dict insertion order; alive[id] = shared by every ingroup genome (a key
 * of sharedKmers, getCandidateKmers.py:190-219), pick[id] = PFB_F_PICK.  Host pointers into pinned buffers owned
 * by the context, valid until its next run.  info = (k - 1) | gc_count << 5 | flags << 11 (PFB_F_GC_OK,
 * PFB_F_TM_OK, PFB_F_REP_OK, PFB_F_PALIN); pos[g * pitch + id] = genome-local padded coordinate of the leftmost
 * base | strand << 31 in the g-th genome of this context, 0xFFFFFFFF where the k-mer is not unique in that genome
 * (alive[id] = none of the rows of any rank says so) -- the contigs of a genome are
 * laid out in the order added, each starting at a multiple of 32, so for a single-contig genome this IS the
 * start; pfb_result_positions / pfb_result_picked decode it to (contig, start). */
typedef struct pfb_idview {
    uint64_t n_ids, pitch, n_genomes;
    uint64_t records_valid;     /* 0: word / tm / info were left on the device (PFB_OPT_FETCH_RECORDS = 0) */
    const uint64_t *word;
    const double   *tm;
    const uint16_t *info;
    const uint32_t *pos;
    const uint8_t  *alive;
    const uint8_t  *pick;
} pfb_idview;

/* lifecycle ------------------------------------------------------------------------------ */
int  pfb_create(pfb_ctx **out, int device, const pfb_params *params);
void pfb_destroy(pfb_ctx *ctx);
const char *pfb_last_error(const pfb_ctx *ctx);   /* ctx may be NULL: error of the last failed pfb_create */
int  pfb_default_params(pfb_params *p);           /* bin/Parameters.py:30-52 + primer3 defaults */

/* input: one call per ingroup genome, in params.ingroupFns order; the first one added is "the first
 * genome" of getCandidateKmers.py:85,108.  `bases` = the contigs concatenated without separators
 * (ASCII; only upper-case A C G T are nucleotides), `contig_off` = n_contigs+1 offsets into it.
 * Replaces the SeqIO.parse + '~'-join at getCandidateKmers.py:89-100.  The buffers are borrowed
 * until pfb_run / pfb_upload returns (pinned memory makes the copy asynchronous). */
int  pfb_add_genome(pfb_ctx *ctx, const uint8_t *bases, uint64_t n_bases,
                    const uint64_t *contig_off, uint32_t n_contigs);
/* the same from the FASTA file itself: `file_bytes` is the file image (host pointer, borrowed like above); the
 * records are found and their sequence lines compacted ON THE DEVICE, with Bio.SeqIO.parse(fn, "fasta")'s
 * rules (getCandidateKmers.py:93,212; biopython's SimpleFastaParser): text before the first line starting
 * with '>' is skipped, '\n' '\r' and ' ' never belong to a sequence, case is preserved.  All genomes of a
 * context must be added the same way.  After pfb_upload / pfb_run, pfb_genome_info and pfb_contig_table
 * report what was found: per record the offset of its '>' in the file (the id is the title up to the first
 * whitespace, SeqRecord.id) and its length. */
int  pfb_add_fasta(pfb_ctx *ctx, const uint8_t *file_bytes, uint64_t n_bytes);
/* the same for a GenBank flat file (the reference's default --format, Bio.SeqIO.parse(fn, "genbank")): a line
 * starting with LOCUS opens a record, the letters of the lines between its ORIGIN line and the next "//" are
 * the sequence, upper-cased.  pfb_contig_table reports the offset of each LOCUS line (the host takes the id
 * from the VERSION / LOCUS lines that follow) and the sequence length. */
int  pfb_add_genbank(pfb_ctx *ctx, const uint8_t *file_bytes, uint64_t n_bytes);
int  pfb_genome_info(pfb_ctx *ctx, uint32_t genome_idx, uint32_t *n_contigs, uint64_t *n_bases);
int  pfb_contig_table(pfb_ctx *ctx, uint32_t genome_idx, uint64_t *header_off, uint64_t *length, uint32_t cap);
int  pfb_clear_genomes(pfb_ctx *ctx);

/* run: pfb_run == pfb_upload + pfb_compute.  pfb_compute can be repeated on resident inputs.
 * Replaces __getSharedKmers (:190-219) and the GC / Tm / repeat part of __evaluateKmersAtOnePosition
 * (:276-320, 376-388) + grouping (:222-247). */
int  pfb_upload(pfb_ctx *ctx);
int  pfb_compute(pfb_ctx *ctx);
int  pfb_run(pfb_ctx *ctx);
/* the two halves of pfb_run's overlap for callers that drive the stages themselves (multi-GPU): the copies
 * are queued on a second stream, pfb_compute_stage(0) consumes each genome as it lands, pfb_upload_wait
 * returns when the host buffers may be reused */
int  pfb_upload_async(pfb_ctx *ctx);
int  pfb_upload_wait(pfb_ctx *ctx);

/* multi-GPU, one process per GPU (replaces the reference's multiprocessing.Pool at getCandidateKmers.py:115-118,
 * 433-436).  Genome 0 is added on every rank -- it defines the key table and the ids, identically everywhere --
 * the other genomes are dealt out; genome 0's own pipeline (candidate windows, its scan) is shared out by
 * position.  Rank 0 makes the id, hands it to the others by any means (128 bytes), every rank joins; from then
 * on pfb_run / pfb_compute are collective calls: inside, the ranks all-gather genome 1's candidate lists,
 * SUM-reduce its packed rows, MIN-reduce one byte per id ("shared by every genome this rank holds") and
 * MAX-reduce one byte per id ("picked in some genome") with NCCL on the context's stream. */
int  pfb_comm_unique_id(uint8_t *id128);
int  pfb_comm_init_rank(pfb_ctx *ctx, int n_ranks, int rank, const uint8_t *id128);
int  pfb_sync(pfb_ctx *ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on (timing with the caller's events) */
int  pfb_stream(pfb_ctx *ctx, void **stream_out);
#define PFB_OPT_SCAN_MODE      0   /* see pfb_set_scan_mode */
#define PFB_OPT_FETCH_RECORDS  1   /* 0: pfb_run leaves the per-id word / Tm / info columns on the device (they are the
                                      same on every rank of a job: ranks > 0 need not move them) */
#define PFB_OPT_SHARD_FIRST    2   /* 0: every rank runs all of genome 0's pipeline itself (default 1) */
#define PFB_OPT_EMULATE_RANKS  3   /* n > 1 (tests): one context plays n ranks of genome 0's shared-out pipeline in turn --
                                      candidate lists, table update, packed rows, polluter pass: everything but NCCL */
int  pfb_set_option(pfb_ctx *ctx, int option, int value);

/* multi-GPU, one process for all GPUs: the same run driven by one host thread over N contexts.  Genomes are
 * added once, in params.ingroupFns order (the first to every GPU, the others round-robin); pfb_multi_locate says
 * where a genome went, pfb_multi_context gives the context whose result calls then apply (genome indices there
 * are the context's own). */
typedef struct pfb_multi pfb_multi;
int  pfb_multi_create(pfb_multi **out, const int *device_ids, int n_devices, const pfb_params *params);
void pfb_multi_destroy(pfb_multi *m);
const char *pfb_multi_last_error(const pfb_multi *m);
int  pfb_multi_add_genome(pfb_multi *m, const uint8_t *bases, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs);
int  pfb_multi_add_fasta(pfb_multi *m, const uint8_t *file_bytes, uint64_t n_bytes);
int  pfb_multi_add_genbank(pfb_multi *m, const uint8_t *file_bytes, uint64_t n_bytes);
int  pfb_multi_run(pfb_multi *m);
int  pfb_multi_size(const pfb_multi *m);
int  pfb_multi_context(pfb_multi *m, int index, pfb_ctx **out);
int  pfb_multi_locate(const pfb_multi *m, uint32_t genome_idx, int *context_index, uint32_t *local_idx);

/* output ---------------------------------------------------------------------------------- */
int  pfb_result_counts(pfb_ctx *ctx, uint64_t *n_records, uint64_t *n_picked);
/* the hot output path: the per-id columns in pinned host memory.  pfb_run / pfb_multi_run copy them out WHILE the
 * run goes on (the records as soon as the ids exist, every genome's places as soon as it has been scanned), so
 * that after the run only two bytes per id are still on their way; after pfb_compute the first result call
 * fetches them.  What __buildOutput (getCandidateKmers.py:442-489) needs is word / k / Tm / GC of the picked ids
 * and their place in every genome: columns, no per-record structs. */
int  pfb_result_by_id(pfb_ctx *ctx, pfb_idview *out);
int  pfb_result_records(pfb_ctx *ctx, pfb_record *out, uint64_t cap);
/* positions of every record (picked_only = 0) or of the picked records only, in record order, in
 * the genome that was added as number `genome_idx` on this context */
int  pfb_result_positions(pfb_ctx *ctx, uint32_t genome_idx, int picked_only, pfb_pos *out, uint64_t cap);
/* the picked records (PFB_F_PICK) compacted, and, for every genome g of this context, start_strand_out[g * cap
 * + i] = leftmost start within the contig | strand << 31 and contig_out[g * cap + i] = contig index of picked
 * record i (contig_out may be NULL).  cap >= n_picked is the row pitch of the two arrays.  Compacted on the host
 * from the per-id columns. */
int  pfb_result_picked(pfb_ctx *ctx, pfb_record *rec_out, uint32_t *start_strand_out, uint32_t *contig_out, uint64_t cap);
int  pfb_host_alloc(void **out, uint64_t n_bytes);    /* cudaHostAlloc */
int  pfb_host_free(void *ptr);
int  pfb_timings(pfb_ctx *ctx, pfb_timing *out);
/* how the scan probes the key bitmap: 0 auto (default), 1 plain L2 probe (k_scan<SCAN>), 2 through the
 * 192 KB shared-memory filter (k_scan_filtered).  Results are identical; tests run both. */
int  pfb_set_scan_mode(pfb_ctx *ctx, int mode);

/* single-oligo helper running the device Tm / GC code on one sequence (known-answer tests,
 * README.md:375-377); returns PFB_EINVAL for non-ACGT input */
int  pfb_oligo_tm(pfb_ctx *ctx, const char *seq, double *tm, double *gc_percent);

/* outgroup screen -------------------------------------------------------------------------
 * Replaces _removeOutgroupPrimers' scan and product arithmetic (bin/removeOutgroupPrimers.py:181-313): the
 * Aho-Corasick automaton over the distinct primer strings of the pairs (:14-31), its run over every outgroup
 * contig on both strands (:34-62, 215-226), the product sizes per pair and contig (:65-122) and the deletion of
 * pairs with a size in params.disallowedLens (:281-287).  A handle lives on one GPU and is not thread-safe.
 * Keys are the DISTINCT primer strings as (2-bit word, length) -- A0 T1 C2 G3, first base most significant, like
 * pfb_record.word; pairs are (key of str(fwd), key of str(rev)) in the order of the pairs dict. */
typedef struct pfb_og pfb_og;
typedef struct pfb_og_hit {       /* one entry of __extractKmerData's lists (:34-62) */
    uint32_t key;                 /* index into the key set */
    uint32_t genome, contig;      /* outgroup genome in pfb_og_add_genome order, contig index within it */
    uint32_t start;               /* '+': end - len + 1 (:51); '-': len(seq) - start_in_rc - 1 = the rightmost (+) coordinate (:54-55) */
    uint32_t strand;              /* 0 '+', 1 '-' */
} pfb_og_hit;
typedef struct pfb_og_product {   /* one (contig, pcrLen) of a pair that is NOT disallowed (:289-291); duplicates possible */
    uint32_t pair, genome, contig;
    int32_t  size;                /* rStart - fStart + 1 > 0 (:80-85) */
} pfb_og_product;
typedef struct pfb_og_timing {
    float    ms_upload, ms_encode, ms_scan, ms_bucket, ms_products, ms_total;   /* CUDA events; total = encode .. products */
    uint64_t n_bases, n_windows, n_hits, n_products;
    uint64_t n_probe_occupied;    /* probes that met an occupied table slot (the others ended on their first load) */
    uint64_t n_launches, n_reruns, table_bytes;
    uint64_t filter_in_smem;      /* 1: the scan tested its windows against the filter in shared memory (k_og_scan_smem) */
    uint64_t filter_bits_per_record;
} pfb_og_timing;
int  pfb_og_create(pfb_og **out, int device);
void pfb_og_destroy(pfb_og *og);
const char *pfb_og_last_error(const pfb_og *og);   /* og may be NULL: error of the last failed pfb_og_create */
int  pfb_og_set_keys(pfb_og *og, const uint64_t *word, const uint8_t *k, uint32_t n_keys);     /* __getAutomaton (:14-31) */
/* disallowed_lo..disallowed_hi inclusive = params.disallowedLens (a range, bin/Parameters.py:678,762,874) */
int  pfb_og_set_pairs(pfb_og *og, const uint32_t *fwd_key, const uint32_t *rev_key, uint32_t n_pairs,
                      int32_t disallowed_lo, int32_t disallowed_hi);
/* one call per outgroup genome (same conventions as pfb_add_genome; an empty contig is allowed); the bytes are
 * str(contig.seq) (:221): only upper-case A C G T can be part of a match */
int  pfb_og_add_genome(pfb_og *og, const uint8_t *bases, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs);
int  pfb_og_clear_genomes(pfb_og *og);
/* 0 auto (default): the window filter lives in shared memory when the key set allows; 1: always in global memory.
 * Results are identical; tests run both. */
int  pfb_og_set_scan_mode(pfb_og *og, int mode);
int  pfb_og_run(pfb_og *og);        /* upload + scan every genome + products */
int  pfb_og_compute(pfb_og *og);    /* the same again on the resident genomes (timing) */
int  pfb_og_counts(pfb_og *og, uint64_t *n_hits, uint64_t *n_products, uint64_t *n_removed);
int  pfb_og_hits(pfb_og *og, pfb_og_hit *out, uint64_t cap);              /* unordered */
int  pfb_og_products(pfb_og *og, pfb_og_product *out, uint64_t cap);      /* unordered */
/* per pair: the first outgroup genome (pfb_og_add_genome order) in which it has a disallowed product size -- the
 * pass of the loop at :255-291 that deletes it -- or 0xFFFFFFFF when the pair is kept */
int  pfb_og_removed_at(pfb_og *og, uint32_t *out, uint64_t cap);
int  pfb_og_timings(pfb_og *og, pfb_og_timing *out);

/* pair search --------------------------------------------------------------------------------
 * Replaces the arithmetic of _getPrimerPairs (bin/getPrimerPairs.py:490-564): binning of every genome's candidates
 * (__binCandidateKmers :39-98), the substring reduction of the first genome's bins (__reduceBinSize :10-36), the
 * bin pairs (__getBinPairs :101-143), the first suitable primer pair of every bin pair (the generator of
 * __getCandidatePrimerPairs :283-341), the look-up of every pair in every other genome (__getAllSharedPrimerPairs
 * :367-464) and its bin numbers there (__updateBinsForUnprocessedGenomes :467-502).  primer3's heterodimer check
 * (:146-178, 330-341) stays the caller's: it filters the proposed pairs, each of which is independent of the others.
 * Candidates are columns: the S candidate k-mers (the same set in every genome, getCandidateKmers.py:442-489) as
 * 2-bit words reading like the first genome's (+) strand, lengths and Tm; per genome the candidates in the order of
 * its dict of lists (contigs in dict order, each list in list order). */
typedef struct pfb_pp pfb_pp;
typedef struct pfb_pp_params {
    int32_t max_bin;          /* params.maxBinSize (:533) */
    int32_t min_primer_len;   /* params.minLen (:536) */
    int32_t min_pcr, max_pcr; /* params.minPcr / maxPcr */
    double  max_tm_diff;      /* params.maxTmDiff (:227) */
} pfb_pp_params;
typedef struct pfb_pp_pair {  /* one pair the generator yields (:297-341), in its order; first-genome values */
    uint32_t fwd, rev;        /* candidate indices: p1 = fwd, p2 = reverse complement of rev (:333) */
    uint32_t contig, fbin, rbin, pcr_len;
} pfb_pp_pair;
#define PFB_PP_VALID 0x80000000u   /* pfb_pp_shared: the pair has a Product in this genome; the low bits are its size */
typedef struct pfb_pp_timing {
    float    ms_bin, ms_pairs, ms_shared, ms_total;    /* CUDA events: binning of all genomes; reduction + bin pairs + primer pairs; cross-genome look-up */
    uint64_t n_candidates, n_genomes, n_bins, n_pairs, n_kept, n_launches;
} pfb_pp_timing;
int  pfb_pp_create(pfb_pp **out, int device);
void pfb_pp_destroy(pfb_pp *pp);
const char *pfb_pp_last_error(const pfb_pp *pp);   /* pp may be NULL: error of the last failed pfb_pp_create */
int  pfb_pp_set_candidates(pfb_pp *pp, const uint64_t *word, const uint8_t *k, const double *tm, uint64_t n);
/* one call per genome, the first genome first: n = the number of candidates; per list position the candidate index,
 * the contig number (dense, in dict order: non-decreasing), min(Primer.start, Primer.end) and the strand (0 '+', 1 '-').
 * The columns are copied to the device inside the call (pfb_pp_set_candidates likewise). */
int  pfb_pp_add_genome(pfb_pp *pp, const uint32_t *id, const uint32_t *contig, const uint32_t *left, const uint8_t *strand, uint64_t n);
int  pfb_pp_clear_genomes(pfb_pp *pp);
int  pfb_pp_run(pfb_pp *pp, const pfb_pp_params *params);
int  pfb_pp_counts(pfb_pp *pp, uint64_t *n_pairs, uint64_t *n_kept, uint64_t *n_bins);   /* kept = valid in every genome (:459-462) */
int  pfb_pp_pairs(pfb_pp *pp, pfb_pp_pair *out, uint64_t cap);
/* out[genome * n_pairs + pair] (genome 0 = the first genome) = PFB_PP_VALID | product size (:455-463), or 0 when the
 * pair has no Product there (other contig / other strand / size out of range).  The Product's bin numbers (:501-502) are
 * pfb_pp_bins(genome)[fwd] and [rev]. */
int  pfb_pp_shared(pfb_pp *pp, uint32_t *out, uint64_t cap);
int  pfb_pp_kept(pfb_pp *pp, uint8_t *out, uint64_t cap);           /* [n_pairs] 1 = a Product in every genome (:459-462) */
/* bin number of every candidate within its contig in one genome; reduced (first genome only, may be NULL): 1 = dropped
 * from its bin by __reduceBinSize */
int  pfb_pp_bins(pfb_pp *pp, uint32_t genome, uint32_t *bin_of, uint8_t *reduced, uint64_t cap);
int  pfb_pp_timings(pfb_pp *pp, pfb_pp_timing *out);

#ifdef __cplusplus
}
#endif
#endif /* PFB200_H */
