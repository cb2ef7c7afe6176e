/* compseed_b200 -- C-ABI of the B200-native SMEM seeding path.
 *
 * This is the drop-in boundary for the seeding hot path of i-xiaohu/CompSeed (a BWA-MEM 0.7.17 fork).
 * The reference has no plugin/FFI interface (it is statically linked C/C++), so the seam is
 * data-shaped: everything upstream of mem_chain() is replaced by the calls below.
 * Plain pointers and sizes only; no torch / CUDA types cross this boundary.
 * Reference paths are relative to the reference tree (i-xiaohu/CompSeed).
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   ----------------------------------------
 *   bwt_t after bwt_restore_bwt/bwt_restore_sa             cs_index_upload / cs_index_load
 *     (FM_index/bwt.c:421-462, bwalib/bwa.c:288)
 *   bwt_occ4 / bwt_2occ4   (FM_index/bwt.c:169,189)        cs_occ4            (unit-level probe)
 *   bwt_extend             (FM_index/bwt.c:262)            cs_extend          (unit-level probe)
 *   bwt_sa                 (FM_index/bwt.c:86)             cs_sa              (also used by the batch path)
 *   mem_collect_intv       (mapping/bwamem.c:218-272)      cs_seed_batch_submit / cs_seed_batch_wait
 *   seeding block of seed_and_extend                        "        (mems[] == aux.match[r], sorted by info)
 *     (mapping/comp_seed.cpp:2255-2302)
 *   seed expansion + bwt_sa                                 "        (rbeg[] == seed[r][*].rbeg, emission order)
 *     (mapping/bwamem.c:386-399, comp_seed.cpp:2306-2346)
 *   BandedPairWiseSW::scalarBandedSWAWrapper / getScores8  cs_bsw_extend      (one ksw_extend2 per sequence pair)
 *     / getScores16 (mapping/bandedSWA.cpp, ksw.c:380)
 *   kt_for workers over reads / 512-read blocks            slots of a cs_ctx_t (pinned buffers, one CUDA
 *     (mapping/bwamem.c:1343, comp_seed.cpp:2541-2548)       stream per slot; submit batch i+1 while waiting on i)
 *
 * Error convention: every int-returning call returns CS_OK (0) or a negative CS_E_* code and sets a
 * thread-local message readable through cs_last_error().  The reference's convention is "fatal"
 * (err_fatal / xassert, bwalib/utils.c:92-124); the host shim turns non-zero into err_fatal.
 * There is no CPU fallback anywhere behind this ABI: without a CUDA device every call fails.
 */
#ifndef COMPSEED_B200_H
#define COMPSEED_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_OK            0
#define CS_E_ARG        -1   /* bad argument */
#define CS_E_CUDA       -2   /* CUDA runtime error (message in cs_last_error) */
#define CS_E_OVERFLOW   -3   /* a result buffer of the ctx is too small for this batch (see cs_last_error) */
#define CS_E_IO         -4   /* index file could not be read */
#define CS_E_NODEVICE   -5   /* no CUDA device: this library has no CPU path */
#define CS_E_STATE      -6   /* slot used out of order (wait without submit, submit on a busy slot) */
#define CS_E_READ_OVERFLOW -7 /* ONE read needs more mems / interval-list entries than the per-read scratch of the ctx holds
                                (max_read_len fixes it: 2*max_read_len+16 mems, max_read_len list entries); larger max_mems /
                                max_seeds do not help -- unlike CS_E_OVERFLOW, where cs_ctx_need says how much would */

/* The fields of bwt_t that the seeding path reads (FM_index/bwt.h:48-60), in the reference's own
 * in-memory layout: bwt[] = 64-byte buckets (4 x u64 checkpoint + 8 x u32 of 16 bases, MSB first),
 * sa[0] == (uint64_t)-1.  The caller keeps ownership; cs_index_upload copies. */
typedef struct {
	uint64_t primary;
	uint64_t L2[5];
	uint64_t seq_len;
	uint64_t bwt_size;      /* uint32 words in bwt[] */
	const uint32_t *bwt;
	int32_t  sa_intv;       /* power of two */
	uint64_t n_sa;
	const uint64_t *sa;
} cs_bwt_view_t;

/* == bwtintv_t (FM_index/bwt.h:62-64): x[0]=k, x[1]=l, x[2]=s, info = (start << 32) | end */
typedef struct { uint64_t x[3]; uint64_t info; } cs_mem_t;

/* The seeding scalars of mem_opt_t (mapping/comp_seed.h:41-73).  split_len is computed by the
 * caller exactly as its own code does: (int)(min_seed_len*split_factor+.499) in float (bwamem.c:223)
 * or (int)(1.0*min_seed_len*split_factor+.499) in double (comp_seed.cpp:2279). */
typedef struct {
	int32_t min_seed_len;   /* -k, default 19  */
	int32_t split_len;      /* from -r, default 28 */
	int32_t split_width;    /* -s, default 10  */
	int32_t max_mem_intv;   /* -y, default 20; 0 disables round 3 */
	int32_t max_occ;        /* -c, default 500 */
} cs_seed_opt_t;

/* Counters with the meaning of the reference's profile (comp_seed.h:158-160, main.cpp:203-214):
 * queries = logical requests, calls = those that touched the FM-index / suffix array in HBM. */
typedef struct {
	uint64_t ext_queries;   /* bwt_extend requests issued by the seeding kernels */
	uint64_t ext_calls;     /* of those, served from the 64-row Occ buckets (the rest: k-mer table / short-circuits) */
	uint64_t sal_queries;   /* SA rows requested (== number of seeds) */
	uint64_t sal_calls;     /* LF steps walked for them (0 with a dense SA) */
} cs_counters_t;

typedef struct {
	uint32_t n_reads;
	uint64_t n_mems, n_seeds;
	/* Pinned host memory owned by the ctx, valid until the slot is resubmitted.  NULL for
	 * cs_seed_batch_wait_device (results stay in HBM). */
	const uint32_t *mem_off;   /* [n_reads+1] */
	const cs_mem_t *mems;      /* [n_mems]  per read sorted ascending by info (duplicates kept) */
	const uint32_t *seed_off;  /* [n_reads+1] */
	const int64_t  *rbeg;      /* [n_seeds] SA[x0 + k*step] in the emission order of bwamem.c:386-399 */
	cs_counters_t counters;
	float kernel_ms[8];        /* CUDA-event durations on the slot's stream: [0] seeding (k_pack_reads + k_seed_fast + k_seed_walk + k_seed +
	                              k_seed_r3), [1] collect, [2] SA-resolve, [3] whole slot incl. copies, [4] passes 1-2 (k_pack_reads +
	                              k_seed_fast + k_seed_walk + k_seed), [5] k_seed_r3 alone, [6] k_pack_reads + k_seed_fast, [7] k_seed_walk */
	uint64_t n_deferred;       /* bwt_smem1a calls the fast kernel handed to the literal kernel */
	/* EXECUTED memory requests (one per lane and load instruction that leaves the SM towards L2: Occ sectors, filter words,
	 * table entries, SA / inverse-SA words, text words), counted in the kernels: [0] k_seed_fast, [1] k_seed_walk,
	 * [2] k_seed, [3] the third-pass kernel, [4] k_sa_resolve, [5] unused.  The roofline of bench.py is built on these. */
	uint64_t gather_requests[6];
	float kernel_ms2[4];       /* [0] k_seed (literal) alone, [1] third-pass kernel alone (it overlaps k_seed_walk / k_seed on the slot's
	                              second stream: kernel_ms[5] is then only what is left of it after k_seed), [2] k_pack_reads, [3] unused */
} cs_result_t;

/* Compact wire format of the results (cs_ctx_config_t.compact_results): what crosses the device-to-host link is 20
 * bytes per mem and 5 per seed position instead of 32 and 8 (on 150-bp reads 225 instead of 355 bytes per read; the link,
 * not the GPU, bounds the host-buffer path).  Lossless: seq_len < 2^37 (cs_index_upload rejects more) and read positions
 * < 2^16 (comp_seed.h:39).  A consumer expands one read at a time where it needs bwtintv_t (cs_cmem_unpack: the kt_for
 * workers do it in parallel, in place of the memcpy they would do anyway), or all at once (cs_compact_expand). */
typedef struct {
	uint32_t x0, x1, x2;    /* low 32 bits of k, l, s */
	uint32_t info;          /* start << 16 | end */
	uint32_t hi;            /* bits 0-4: k >> 32, 5-9: l >> 32, 10-14: s >> 32 */
} cs_cmem_t;

static inline void cs_cmem_unpack(const cs_cmem_t *c, cs_mem_t *m)
{
	m->x[0] = (uint64_t)c->x0 | ((uint64_t)(c->hi & 31u) << 32);
	m->x[1] = (uint64_t)c->x1 | ((uint64_t)((c->hi >> 5) & 31u) << 32);
	m->x[2] = (uint64_t)c->x2 | ((uint64_t)((c->hi >> 10) & 31u) << 32);
	m->info = ((uint64_t)(c->info >> 16) << 32) | (uint64_t)(c->info & 0xffffu);
}
/* seed position i: 40 bits, sign-extended (bwt_sa of row 0 is (uint64_t)-1, FM_index/bwt.c:83,93) */
static inline int64_t cs_crbeg(const uint32_t *lo, const uint8_t *hi, uint64_t i)
{ return (int64_t)(((uint64_t)lo[i] | ((uint64_t)hi[i] << 32)) << 24) >> 24; }

typedef struct {
	uint32_t n_reads;
	uint64_t n_mems, n_seeds;
	const uint32_t *mem_off;   /* [n_reads+1] */
	const cs_cmem_t *cmems;    /* [n_mems] per read sorted ascending by info */
	const uint32_t *seed_off;  /* [n_reads+1] */
	const uint32_t *rbeg_lo;   /* [n_seeds] */
	const uint8_t  *rbeg_hi;   /* [n_seeds] */
	cs_counters_t counters;
} cs_compact_result_t;

typedef struct cs_index cs_index_t;
typedef struct cs_ctx cs_ctx_t;

/* Result-neutral structures built next to the index.  Every field: -1 = the default for this
 * index size.  None of them changes a result; they are here so that a host (and the tests) can switch each one off. */
typedef struct {
	int32_t kmer_table_depth;  /* top-of-search table of every string of <= this many bases (default: up to 13); 0 = none */
	int32_t prune_k;           /* K of the K-mer occurrence filter (default ceil(log4 seq_len)+2, at most 19); 0 = none */
	int32_t isa_intv;          /* sampling of the inverse SA used by the unique-match paths (power of two, default 2); 0 = no
	                              2-bit text / inverse SA (the fast kernels are then not used) */
	int32_t repeat_lengths;    /* 1 (default -1 = 1): one byte per text position, the length of the longest repeat starting there
	                              (second-pass calls inside a one-occurrence SMEM then read neither the FM-index nor the filter); 0 = none.
	                              Needs the 2-bit text (isa_intv > 0) */
} cs_index_config_t;
void cs_index_config_default(cs_index_config_t *cfg);

/* Execution switches of a ctx.  -1 / 0 as documented per field; none of them changes a result. */
typedef struct {
	int32_t use_fast;          /* 1 (default -1 = 1): k_seed_fast / k_seed_walk where the index allows; 0: the literal kernel alone */
	int32_t use_r3_fast;       /* 1 (default): text-assisted third pass; 0: k_seed_r3 */
	int32_t defer_cap;         /* capacity of the deferred-call queue (default -1: 2*max_reads + 4096); a batch that overflows it
	                              is rerun through the literal kernel alone */
	int32_t lit_ctas_per_sm;   /* cap on resident CTAs of the literal kernel in call mode (default -1: no cap) */
	int32_t prefetch_results;  /* 1: result copies of finished slots are enqueued while the host waits on another (default 0) */
	int32_t l2_persist_mb;     /* > 0: an L2 access-policy window (persisting) of this many MB over the top of the K-mer table on
	                              every slot stream (default 0: none; profiles/ has the measurement) */
	int32_t overlap_streams;   /* 1: the third-pass kernel of a batch runs on a forked stream next to k_seed_walk / k_seed, on which it
	                              does not depend.  Default 0: measured neutral on one B200 (every kernel of the batch fills the GPU: what
	                              the third pass gains, k_seed loses; profiles/r02_variants.json), and serial kernels time cleanly */
	int32_t compact_results;   /* 1: the batches also produce the compact wire format (cs_seed_batch_wait_compact); default 0 */
	int32_t batch_order;       /* how the kernels of batches submitted to different slots of the ctx share the GPU.  1 (default -1 = 1): the
	                              first kernel of a batch waits for the last kernel of the batch submitted before it, so batches finish one
	                              after the other and the result copy of one overlaps the kernels of the next.  0: no ordering -- every
	                              kernel fills the GPU, so the batches in flight take turns kernel by kernel, all finish together, their
	                              result copies queue up and the GPU idles meanwhile (profiles/r02_pipeline_timeline.md) */
} cs_ctx_config_t;
void cs_ctx_config_default(cs_ctx_config_t *cfg);

const char *cs_last_error(void);
int cs_last_error_code(void);   /* the CS_E_* code that goes with it (for the calls that return a handle, not a code) */
int cs_device_count(void);

/* --- index ------------------------------------------------------------------------------------
 * The device copy is re-laid-out once at upload: 32-byte buckets of 64 BWT rows (128-bit base words
 * + 3 x 40-bit checkpoint), read with one 256-bit load.  dense_sa_intv: 0 keeps the sampling of the
 * input; a power of two < sa_intv re-samples the suffix array on the device (bwt_sa(k) does not
 * depend on the sampling, FM_index/bwt.c:86-96), 1 = full suffix array. */
cs_index_t *cs_index_upload(const cs_bwt_view_t *bwt, int device, int dense_sa_intv);
cs_index_t *cs_index_upload_ex(const cs_bwt_view_t *bwt, int device, int dense_sa_intv, const cs_index_config_t *cfg /* NULL = defaults */);
/* Reads P.bwt and P.sa exactly as bwt_restore_bwt / bwt_restore_sa do (FM_index/bwt.c:421-462). */
cs_index_t *cs_index_load(const char *prefix, int device, int dense_sa_intv);
cs_index_t *cs_index_load_ex(const char *prefix, int device, int dense_sa_intv, const cs_index_config_t *cfg);
/* Writes P.bwt and P.sa exactly as bwt_dump_bwt / bwt_dump_sa do (FM_index/bwt.c:385-407), at sampling sa_intv (the
 * reference writes 32): an index built by cs_index_build can be loaded by the reference's bwt_restore_bwt/sa. */
int cs_index_write(const cs_index_t *idx, const char *prefix, int sa_intv);
/* Builds the FM-index of fwd+revcomp(fwd) on the device (fwd: l_pac nt4 codes 0..3 in host memory).
 * Same BWT / Occ / SA as bwaidx (FM_index/index_main.c:257-325); the SA is kept at sa_intv rows. */
cs_index_t *cs_index_build(const uint8_t *fwd, uint64_t l_pac, int device, int sa_intv);
cs_index_t *cs_index_build_ex(const uint8_t *fwd, uint64_t l_pac, int device, int sa_intv, const cs_index_config_t *cfg);
/* Self-check of a device index (any origin), by the definitions rather than by comparison with another builder:
 *   out[0] rows checked for "suffix SA[r-1] < suffix SA[r]" through the 2-bit text, out[1] violations
 *   out[2] rows checked for "BWT[r] == T[SA[r]-1]",                               out[3] violations
 *   out[4] text positions hit twice or never by SA (permutation test, full),      out[5] Occ checkpoints that are not
 *          the running base counts of the BWT / L2 / primary inconsistencies
 *   out[6] inverse-SA samples checked for SA[ISA[p]] == p,                          out[7] violations
 *   out[8] K-mers checked: filter count == min(3, occurrences by backward search), out[9] violations
 *   out[10] top-of-search entries checked against bwt_extend from scratch,        out[11] violations
 *   out[12] text bases compared with fwd / revcomp(fwd) (when fwd != NULL),        out[13] mismatches
 *   out[14] repeat lengths checked (T[p, p+R) occurs twice, one base more once),  out[15] violations
 * Needs the dense SA and the 2-bit text.  stride: every stride-th row / sample (1 = all).  Returns CS_OK when it ran;
 * the caller looks at the violation counts. */
int cs_index_verify(const cs_index_t *idx, const uint8_t *fwd, uint64_t l_pac, uint32_t stride, uint64_t out[16]);
/* Copies the index back in the reference layout (for a host-side consumer such as the CPU baseline).
 * Call once with bwt == NULL to get sizes in *view, then with caller-allocated arrays of
 * view->bwt_size uint32 and view->n_sa uint64 (at the reference sampling out_sa_intv, e.g. 32). */
int cs_index_download(const cs_index_t *idx, cs_bwt_view_t *view, uint32_t *bwt, uint64_t *sa, int out_sa_intv);
int cs_index_info(const cs_index_t *idx, cs_bwt_view_t *view /* pointers set to NULL */, uint64_t *device_bytes);
void cs_index_free(cs_index_t *idx);

/* unit-level probes: n queries from host arrays, answered by the same device functions the batch
 * kernels use.  cnt: n*4, ik: n*3 (x0,x1,x2), ok: n*4*3, is_back: n. */
int cs_occ4(const cs_index_t *idx, uint32_t n, const uint64_t *k, uint64_t *cnt);
int cs_extend(const cs_index_t *idx, uint32_t n, const uint64_t *ik, const int32_t *is_back, uint64_t *ok);
int cs_sa(const cs_index_t *idx, uint32_t n, const uint64_t *k, uint64_t *out);

/* --- batches ----------------------------------------------------------------------------------
 * A ctx owns n_slots independent slots (stream + pinned host buffers + device buffers).  One
 * submitting thread per ctx; the index handle may be shared read-only between ctxs.
 * max_mems / max_seeds: result capacities per slot (0: 16 / 32 per read). */
cs_ctx_t *cs_ctx_create(const cs_index_t *idx, uint32_t max_reads, uint64_t max_bases, uint32_t max_read_len,
                        uint64_t max_mems, uint64_t max_seeds, int n_slots);
cs_ctx_t *cs_ctx_create_ex(const cs_index_t *idx, uint32_t max_reads, uint64_t max_bases, uint32_t max_read_len,
                           uint64_t max_mems, uint64_t max_seeds, int n_slots, const cs_ctx_config_t *cfg /* NULL = defaults */);
void cs_ctx_free(cs_ctx_t *ctx);
/* After CS_E_OVERFLOW on a slot: the capacities this batch needs (pass them to a new ctx and redo the batch once). */
int cs_ctx_need(const cs_ctx_t *ctx, int slot, uint64_t *need_mems, uint64_t *need_seeds);
/* Kernels launched on behalf of this ctx so far (ours and the two cub scans per batch), counted at the launch sites. */
uint64_t cs_ctx_launches(const cs_ctx_t *ctx);

/* bases: nt4 codes (0..3, anything > 3 is ambiguous) of all reads concatenated, converted as
 * comp_seed.cpp:2258-2260 does; offsets: n_reads+1.  Copies into the slot's pinned buffer, then
 * enqueues H2D + kernels on the slot's stream and returns without waiting. */
int cs_seed_batch_submit(cs_ctx_t *ctx, int slot, uint32_t n_reads, const uint8_t *bases, const uint32_t *offsets,
                         const cs_seed_opt_t *opt);
/* The same with the reads already packed by the caller (57 instead of 150 bytes per 150-bp read
 * cross the host-to-device link).  Layout: read r owns the words [(offsets[r] >> 5) + 2r, … + (len_r >> 5) + 2);
 * word w of a read holds its bases 32w .. 32w+31, base j at bits 2j of packed[] (0..3) and at bit j of nmask[]
 * (1 = ambiguous, i.e. a code > 3, or past the end of the read; the packed bits of such a base are 0).
 * cs_packed_words gives the length of both arrays.  bwa.c:78-111 / main.cpp:36-58 is where a host would pack. */
uint64_t cs_packed_words(uint32_t n_reads, const uint32_t *offsets);
/* Host-side packer into that layout (plain C++ on n_threads host threads; no device involved): what a reader does once per
 * read where it converts ASCII to nt4 today.  packed / nmask: cs_packed_words(n_reads, offsets) entries each. */
int cs_pack_reads_host(uint32_t n_reads, const uint8_t *bases, const uint32_t *offsets, uint64_t *packed, uint32_t *nmask, int n_threads);
int cs_seed_batch_submit_packed(cs_ctx_t *ctx, int slot, uint32_t n_reads, const uint64_t *packed, const uint32_t *nmask,
                                const uint32_t *offsets, const cs_seed_opt_t *opt);
/* Waits for the slot, copies the results to its pinned host buffers and fills *out. */
int cs_seed_batch_wait(cs_ctx_t *ctx, int slot, cs_result_t *out);

/* The same batch in the compact wire format (the ctx must have been created with compact_results = 1). */
int cs_seed_batch_wait_compact(cs_ctx_t *ctx, int slot, cs_compact_result_t *out);
/* Expands a compact result into plain arrays (mems: n_mems entries, rbeg: n_seeds) on n_threads host threads. */
int cs_compact_expand(const cs_compact_result_t *res, cs_mem_t *mems, int64_t *rbeg, int n_threads);

/* --- chaining on the device ---------------------------------------------------------------
 * mem_chain + mem_chain_flt (mapping/bwamem.c:359-497 == mapping/comp_seed.cpp:241-354) run on the mems and seed positions
 * of a batch where the seeding kernels left them, so that only the filtered chains cross the device-to-host link
 * (about 45 bytes per 150-bp read instead of 225).  Bit-exact, including the order-dependent parts: the B-tree of chains
 * (cstl/kbtree.h, t = 5), ks_introsort by weight (cstl/ksort.h) and the pairwise overlap filter. */
typedef struct {             /* the fields of bntseq_t that bns_intv2rid reads (FM_index/bntseq.h:41-64, bntseq.c:354-378) */
	int64_t l_pac;
	int32_t n_seqs;
	const int64_t *offset;   /* [n_seqs] anns[i].offset */
	const uint8_t *is_alt;   /* [n_seqs] anns[i].is_alt != 0, or NULL */
} cs_bns_view_t;

typedef struct {             /* the chaining scalars of mem_opt_t (mapping/comp_seed.h:41-73); defaults of mem_opt_init in () */
	int32_t w;                /* band width (100) */
	int32_t max_chain_gap;    /* (10000) */
	int32_t min_chain_weight; /* (0) */
	int32_t max_chain_extend; /* (1 << 30) */
	float mask_level;         /* (0.50) */
	float drop_ratio;         /* (0.50) */
} cs_chain_opt_t;

typedef struct {
	int32_t rid;              /* mem_chain_t.rid */
	uint32_t w_kept;          /* w : 29 | kept : 2 | is_alt : 1, the bit-field of mem_chain_t (bwamem.c:286) */
	uint32_t n;               /* seeds in the chain */
	uint32_t l_rep;           /* bases of the read covered by repetitive seeds: frac_rep = (float)l_rep / l_seq (bwamem.c:377-385,424) */
} cs_chain_t;                /* pos == rbeg of the chain's first seed */

typedef struct {
	uint32_t n_reads;
	uint64_t n_chains, n_cseeds;
	const uint32_t *chain_off;   /* [n_reads+1] the chains mem_chain_flt keeps, in its output order */
	const uint32_t *cseed_off;   /* [n_reads+1] first chain seed of each read */
	const cs_chain_t *chains;    /* [n_chains] */
	const uint32_t *rbeg_lo; const uint8_t *rbeg_hi;   /* [n_cseeds] seeds of chain after chain, in chain order (cs_crbeg) */
	const uint16_t *qbeg, *len;  /* [n_cseeds] mem_seed_t.qbeg / .len (score == len) */
} cs_chain_result_t;

/* Every batch submitted after this call is also chained (bns == NULL: switched off again).  Copies the contig table. */
int cs_ctx_set_chaining(cs_ctx_t *ctx, const cs_bns_view_t *bns, const cs_chain_opt_t *opt);
/* Waits for the slot and fetches ONLY the chains (the mems / seed positions stay on the device; cs_seed_batch_fetch still gets them). */
int cs_seed_batch_wait_chains(cs_ctx_t *ctx, int slot, cs_chain_result_t *out);

/* --- banded Smith-Waterman extension ----------------------------------------------
 * Replaces the batch calls of the reference's extension stage, BandedPairWiseSW::scalarBandedSWAWrapper / getScores8 / getScores16
 * (mapping/bandedSWA.cpp:242-260 and the SIMD twins; called from mem_chain2aln_across_reads_V2, comp_seed.cpp:1722-2074), i.e. one
 * ksw_extend2 (bwalib/ksw.c:380-479 == scalarBandedSWA, bandedSWA.cpp:118-237) per sequence pair: score, qle, tle, gtle, gscore,
 * max_off, bit for bit -- including the band that follows the non-zero cells from row to row (ksw.c:463-468), the z-drop test
 * (:456-462) and the cap on w (:401-408).  cs_seqpair_t IS the reference's SeqPair (bandedSWA.h:91-99): the caller fills idr / idq
 * (offsets of the target / the query of the pair in seqBufRef / seqBufQer), len1 (target), len2 (query) and h0; the call fills
 * score .. max_off and leaves the other fields alone.  Sequences are nt4 codes (0-3, 4 = N), one byte per base. */
typedef struct {
	int32_t idr, idq, id;
	int32_t len1, len2;
	int32_t h0;
	int32_t seqid, regid;
	int32_t score, tle, gtle, qle;
	int32_t gscore, max_off;
} cs_seqpair_t;

typedef struct {              /* the constructor arguments of BandedPairWiseSW (bandedSWA.cpp:48-58) */
	int32_t o_del, e_del, o_ins, e_ins;   /* mem_opt_t.o_del .. e_ins (6, 1, 6, 1) */
	int32_t zdrop;                        /* mem_opt_t.zdrop (100) */
	int32_t end_bonus;                    /* pen_clip5 for the left extensions, pen_clip3 for the right ones (5) */
	int8_t mat[25];                       /* mem_opt_t.mat: 5 x 5 scores, mat[target * 5 + query] (bwa_fill_scmat, bwalib/bwa.c:419) */
} cs_bsw_opt_t;

typedef struct cs_bsw cs_bsw_t;
/* Device buffers and page-locked staging for batches of up to max_pairs pairs whose sequences take up to max_ref_bytes /
 * max_qer_bytes (grown on demand).  max_qlen: longest query the DP rows are sized for (also grown on demand). */
cs_bsw_t *cs_bsw_create(int device, uint32_t max_pairs, uint64_t max_ref_bytes, uint64_t max_qer_bytes, uint32_t max_qlen);
void cs_bsw_free(cs_bsw_t *b);
/* scalarBandedSWAWrapper(pairs, seqBufRef, seqBufQer, n_pairs, nthreads, w): blocking; host buffers in, pairs[] updated in place.
 * ref_bytes / qer_bytes: how much of the two buffers the pairs refer to (max over pairs of idr + len1 / idq + len2). */
int cs_bsw_extend(cs_bsw_t *b, cs_seqpair_t *pairs, const uint8_t *seq_buf_ref, uint64_t ref_bytes, const uint8_t *seq_buf_qer, uint64_t qer_bytes,
                  uint32_t n_pairs, int32_t w, const cs_bsw_opt_t *opt);
/* The same on inputs already staged on the device by cs_bsw_stage (benchmarks: inputs resident in HBM); results stay on the device
 * until cs_bsw_fetch.  *kernel_ms: CUDA-event time of the extension kernel; *cells: DP cells it computed. */
int cs_bsw_stage(cs_bsw_t *b, const cs_seqpair_t *pairs, const uint8_t *seq_buf_ref, uint64_t ref_bytes, const uint8_t *seq_buf_qer, uint64_t qer_bytes, uint32_t n_pairs);
int cs_bsw_run_staged(cs_bsw_t *b, int32_t w, const cs_bsw_opt_t *opt, float *kernel_ms, uint64_t *cells);
int cs_bsw_fetch(cs_bsw_t *b, cs_seqpair_t *pairs);
uint64_t cs_bsw_launches(const cs_bsw_t *b);
/* Tuning: resident 128-thread CTAs of the extension kernel per SM when its DP rows live in the HBM scratch (default 8); changes no result. */
int cs_bsw_set_ctas_per_sm(cs_bsw_t *b, int ctas_per_sm);
/* Tuning: 1 (default) keeps the DP rows in shared memory whenever the longest query of the batch lets two CTAs of 32 threads share an
 * SM (queries up to ~700 bases with 16-bit cells); 0 always uses the HBM scratch.  Changes no result. */
int cs_bsw_set_rows_in_smem(cs_bsw_t *b, int on);

/* --- multi-device pipeline ----------------------------------------------------------------------
 * Replaces kt_for(opt->n_threads, worker1 / seed_and_extend) over the reads of a -K batch for the seeding part
 * (mapping/bwamem.c:1343, comp_seed.cpp:2541-2548), and -- with two read sets in flight -- the overlap kt_pipeline gives
 * between batch i+1 and batch i (fastmap.c:76-140, kthread.c:95-107).  One index replica and one ctx per device; the reads
 * of a set are split into one contiguous block per device, in input order, each a multiple of 512 reads (BATCH_SIZE,
 * comp_seed.h:36); one host thread per device pipelines its block in batches through the slots of its ctx.  Results arrive
 * in the compact wire format, by DMA, in page-locked arrays owned by the cs_multi_t; nothing is exchanged between devices
 * and no host thread copies a result: gathering in input order is the accessor cs_multi_read. */
#define CS_MULTI_MAX_DEV 16
typedef struct cs_multi cs_multi_t;

typedef struct {            /* the reads [r0, r1) of a set, seeded by one device in batches of batch_reads */
	uint64_t r0, r1;
	uint32_t batch_reads, n_batches;
	const uint64_t *mem_base, *seed_base;   /* [n_batches+1] where a batch's mems / seeds start in the arrays below */
	const uint32_t *mem_off, *seed_off;     /* [n_batches][batch_reads+1] offsets inside the batch */
	const cs_cmem_t *cmems;                 /* mems of the block, reads in input order, per read sorted by info */
	const uint32_t *rbeg_lo; const uint8_t *rbeg_hi;
	int device;
	/* with cs_multi_set_chaining the block holds CHAINS instead: mem_base / mem_off index chains[], seed_base / seed_off
	 * index the chain seeds (rbeg_lo, rbeg_hi, qbeg, len); cmems is NULL */
	const cs_chain_t *chains;
	const uint16_t *qbeg, *len;
} cs_block_t;

typedef struct {
	uint64_t n_reads, n_mems, n_seeds;
	int n_blocks;
	const cs_block_t *blocks;   /* in input order; valid until the set is submitted again */
	cs_counters_t counters;
	double seconds;             /* submit to the last device finishing */
	double host_s[3];           /* where the slowest device's host thread spent that time: [0] submitting batches (offsets, enqueueing
	                               copies and kernels), [1] polling, checking finished batches and enqueueing their result copies,
	                               [2] idle (nothing was ready) */
	double gpu_ms[4];           /* CUDA-event times summed over the batches of the busiest device (batches overlap, so the sum may
	                               exceed `seconds`): [0] submit -> first kernel (input copy, and waiting for the GPU), [1] seeding kernels,
	                               [2] collect + SA resolution (+ chaining, compaction), [3] last kernel -> results on the host */
} cs_multi_result_t;

/* where read r of the set is: its mems cm[0..n_mems) and the index s0 of its first seed position in (lo, hi) */
static inline void cs_multi_read(const cs_multi_result_t *res, uint64_t r, const cs_cmem_t **cm, uint32_t *n_mems,
                                 const uint32_t **lo, const uint8_t **hi, uint64_t *s0, uint32_t *n_seeds)
{
	int k = 0;
	while (k + 1 < res->n_blocks && r >= res->blocks[k].r1) ++k;
	const cs_block_t *b = &res->blocks[k];
	const uint64_t lr = r - b->r0, bi = lr / b->batch_reads, i = bi * (b->batch_reads + 1ull) + lr % b->batch_reads;
	*cm = b->cmems + b->mem_base[bi] + b->mem_off[i]; *n_mems = b->mem_off[i + 1] - b->mem_off[i];
	*lo = b->rbeg_lo; *hi = b->rbeg_hi; *s0 = b->seed_base[bi] + b->seed_off[i]; *n_seeds = b->seed_off[i + 1] - b->seed_off[i];
}

/* the chains of read r (after cs_multi_set_chaining): ch[0..n_chains), their seeds one chain after the other from index s0 */
static inline void cs_multi_read_chains(const cs_multi_result_t *res, uint64_t r, const cs_chain_t **ch, uint32_t *n_chains,
                                        const cs_block_t **blk, uint64_t *s0)
{
	int k = 0;
	while (k + 1 < res->n_blocks && r >= res->blocks[k].r1) ++k;
	const cs_block_t *b = &res->blocks[k];
	const uint64_t lr = r - b->r0, bi = lr / b->batch_reads, i = bi * (b->batch_reads + 1ull) + lr % b->batch_reads;
	*ch = b->chains + b->mem_base[bi] + b->mem_off[i]; *n_chains = b->mem_off[i + 1] - b->mem_off[i];
	*blk = b; *s0 = b->seed_base[bi] + b->seed_off[i];
}

/* idx[k]: the index replica on the k-th device to use (cs_index_replicate copies one device-to-device over NVLink).
 * mems_per_read / seeds_per_read: sizing estimates (0: 16 / 32); buffers grow when a batch needs more. */
cs_multi_t *cs_multi_create(cs_index_t *const *idx, int n_dev, uint32_t batch_reads, uint32_t max_read_len, int n_slots,
                            uint32_t mems_per_read, uint32_t seeds_per_read, const cs_ctx_config_t *cfg);
void cs_multi_free(cs_multi_t *m);
cs_index_t *cs_index_replicate(const cs_index_t *src, int device);
/* Starts seeding a read set and returns at once.  set: 0 or 1 (two sets may be in flight: submit batch i+1 while the host
 * consumes batch i).  offsets: n_reads+1, 64-bit.  bases (nt4 codes) / packed+nmask must stay valid until cs_multi_wait;
 * page-locked memory (cs_host_register) is read by DMA without a staging copy.  The packed form is the set-global layout
 * cs_pack_reads_host64 writes: read r owns the words [(offsets[r] >> 5) + 2r, ... + (len_r >> 5) + 2). */
int cs_multi_submit(cs_multi_t *m, int set, uint64_t n_reads, const uint8_t *bases, const uint64_t *offsets, const cs_seed_opt_t *opt);
int cs_multi_submit_packed(cs_multi_t *m, int set, uint64_t n_reads, const uint64_t *packed, const uint32_t *nmask,
                           const uint64_t *offsets, const cs_seed_opt_t *opt);
int cs_multi_wait(cs_multi_t *m, int set, cs_multi_result_t *out);
/* From the next set on, the devices also chain (cs_ctx_set_chaining) and the results are chains (cs_multi_read_chains); bns ==
 * NULL switches back to mems + seed positions.  No set may be in flight. */
int cs_multi_set_chaining(cs_multi_t *m, const cs_bns_view_t *bns, const cs_chain_opt_t *opt);
int cs_pack_reads_host64(uint64_t n_reads, const uint8_t *bases, const uint64_t *offsets, uint64_t *packed, uint32_t *nmask, int n_threads);
/* Flat arrays in input order from a multi result (mem_off / seed_off: n_reads+1 each; mems / rbeg may be NULL). */
int cs_multi_gather(const cs_multi_result_t *res, uint64_t *mem_off, cs_mem_t *mems, uint64_t *seed_off, int64_t *rbeg, int n_threads);
uint64_t cs_multi_launches(const cs_multi_t *m);
/* Diagnostics: the timeline of the batches device k ran for the last finished run of `set`, 8 floats per batch, milliseconds since the
 * device's ctx was created: GPU clock [0] input copy enqueued, [1] first kernel starts, [2] last kernel done, [3] result copies start,
 * [4] results on the host; host clock [5] submit call, [6] kernels seen finished (result copies enqueued), [7] results seen on the host.
 * Returns the number of batches; at most cap_batches rows are written to out (may be NULL). */
uint32_t cs_multi_trace(const cs_multi_t *m, int set, int k, float *out, uint32_t cap_batches);
/* The block of reads [*r0, *r1) the k-th of n_dev devices gets from a set of n_reads (what cs_multi_submit applies): contiguous,
 * in input order, whole multiples of 512 reads (comp_seed.h:36) except the last.  Plain host arithmetic, no device involved. */
void cs_multi_block_bounds(uint64_t n_reads, int n_dev, int k, uint64_t *r0, uint64_t *r1);

/* Page-locks a caller-owned buffer (e.g. the read buffer of the batch loop).  cs_seed_batch_submit
 * then DMAs straight out of it instead of staging through the slot's pinned buffer; the caller must
 * leave the bases of a submitted batch untouched until the slot is waited on. */
int cs_host_register(void *ptr, size_t bytes);
int cs_host_unregister(void *ptr);

/* Device-resident variant (inputs already in HBM, results left in HBM): stage once, run many. */
int cs_seed_batch_stage(cs_ctx_t *ctx, int slot, uint32_t n_reads, const uint8_t *bases, const uint32_t *offsets);
int cs_seed_batch_run_staged(cs_ctx_t *ctx, int slot, const cs_seed_opt_t *opt);
int cs_seed_batch_wait_device(cs_ctx_t *ctx, int slot, cs_result_t *out);
/* Copies the results of the last run on this slot to the pinned host buffers (after *_wait_device). */
int cs_seed_batch_fetch(cs_ctx_t *ctx, int slot, cs_result_t *out);

/* --- measurement helpers ----------------------------------------------------------------------
 * Random-gather probe: the "HBM random sector" roofline denominator.
 * n_loads independent uniformly random `granule`-byte aligned loads (32 or 64) over table_bytes. */
int cs_probe_random_gather(int device, uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iters,
                           double *gbytes_per_s, double *gloads_per_s);
/* Same with `unroll` (1 or 4) independent loads in flight per thread and an optional
 * cudaLimitMaxL2FetchGranularity (0 = leave as is, else 32 / 64 / 128). */
int cs_probe_random_gather_ex(int device, uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iters, int unroll,
                              int l2_fetch_granularity, double *gbytes_per_s, double *gloads_per_s);
/* The same probe over the index's OWN arrays (Occ buckets, suffix array, occurrence filter, text, inverse SA, table:
 * each load picks an array with probability proportional to its size): the random-sector peak of exactly the memory
 * the seeding kernels gather from, TLB reach included.  granule 32. */
int cs_probe_index_gather(const cs_index_t *idx, uint64_t n_loads, int iters, int unroll, double *gbytes_per_s, double *gloads_per_s);
/* Raw device counters of the last finished run on a slot: [0] ext queries, [1] ext calls, [2] extends whose
 * k and l needed two sectors, [3] occurrence-filter probes, [4..19] event counters of a -DCS_STATS build (else 0), [20] reads the fast
 * kernel deferred to the literal kernel, [21] CTAs of k_seed_fast (0 = not available), [22] CTAs of k_seed, [23] calls the literal kernel ran,
 * [24..39] event counters of k_seed_fast in a -DCS_STATS build. */
int cs_debug_stats(cs_ctx_t *ctx, int slot, uint64_t out[40]);
/* Writes a buffer larger than L2 (flush between timed iterations). */
int cs_flush_l2(int device);

#ifdef __cplusplus
}
#endif
#endif
