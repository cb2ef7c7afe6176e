// Kernels of the seeding path (sm_100a).  Host-side launch code is in cs_api.cu.
#pragma once
#include "cs_device.cuh"
#include "../../include/compseed_b200.h"

#define CS_SEED_BLOCK   256     // threads per CTA of the seeding kernel
#ifndef CS_LIST_SMEM
#define CS_LIST_SMEM    16      // interval-list entries per thread kept in shared memory (the rest spills to HBM).  16 entries and 2 CTAs per
#endif                          // SM instead of 8 and 3: +10 % on the repeat-rich configuration, neutral on cfg2 (profiles/r02_variants.json)
#ifndef CS_READ_SMEM
#define CS_READ_SMEM    9       // 32-base words of the read in flight kept in shared memory (reads <= 256 bases + pad word)
#endif
// dynamic shared memory of k_seed: interval lists [entry][thread] + packed reads [word][thread] + N masks
#define CS_SEED_SMEM_BYTES ((size_t)CS_SEED_BLOCK * (CS_LIST_SMEM * 16 + CS_READ_SMEM * 12))
#ifndef CS_SEED_MINBLOCKS
#define CS_SEED_MINBLOCKS 2     // CTAs per SM the register allocation of k_seed is tuned for
#endif

#define CS_FAST_BLOCK   256     // threads per CTA of k_seed_fast
#ifndef CS_FAST_MINBLOCKS
#define CS_FAST_MINBLOCKS 4
#endif
#ifndef CS_WALK_MINBLOCKS
#define CS_WALK_MINBLOCKS CS_FAST_MINBLOCKS   // CTAs per SM the register allocation of k_seed_walk is tuned for
#endif
#ifndef CS_WALK_MAX
#define CS_WALK_MAX 6           // short forward matches a call may have to be walked by k_seed_walk (else k_seed)
#endif
#ifndef CS_FWD_WIN
#define CS_FWD_WIN 1            // 32-base text windows k_seed_fast fetches at once when it follows a unique match forward.  Measured on
#endif                          // cfg2 (4 M reads, profiles/r02_variants.json): 1: 12.34 ms, 2: 12.52 ms, 4: 13.30 ms -- the extra loads and registers cost more than the shorter dependent chain saves
#ifndef CS_WALK_L
#define CS_WALK_L 1             // calls whose longest forward match has >= K bases but several occurrences go to k_seed_walk (L first), not to k_seed
#endif
#ifndef CS_SPEC_DIAG
#define CS_SPEC_DIAG 0          // k_seed_fast: first-pass calls tried first on the diagonal of the read's last one-occurrence SMEM (needs the repeat lengths).
#endif                          // Correct (tests/test_seed_emul.py runs it), but measured slower on cfg2 (4 M reads: 9.19 ms with, 8.70 ms without,
                                // profiles/r02_variants.json): it is taken for 0.67 calls per read and fails for 0.28, and a failed try costs three gathers
                                // and one more dependent round trip, which the 2.5 requests per read it saves do not pay for
#define CS_FAST_SMEM_BYTES ((size_t)CS_FAST_BLOCK * (CS_READ_SMEM * 12))

struct SeedArgs {
	const uint8_t *bases;       // nt4 codes, concatenated
	const uint32_t *off;        // n_reads + 1
	uint32_t n_reads;
	cs_seed_opt_t opt;
	const uint64_t *packed;     // 2-bit packed reads (k_pack_reads): read r starts at word ((off_bias + off[r]) >> 5) + 2r
	uint32_t off_bias;          // 0..31: the batch is a slice of a larger packed set that starts off_bias bases into its first word
	const uint32_t *nmask;      // ambiguity / end-of-read mask, same word indexing
	// scratch
	uint32_t *next_read;        // work counters: [0] k_seed, [1] k_seed_r3, [2] k_seed_fast, [3] k_seed_walk
	// calls k_seed_fast hands on: {read, y, min_intv, previous deferred call of the same read or ~0} with
	// y = pivot (bits 0-15) | pass (16-17) | d low 5 bits (18-22) | d high 3 bits (24-26) | L-first (27) | walk (31);
	// walk = 1: for k_seed_walk, with the filter bits of its short matches in defer_bits[] (bits 18+ there: 1 + start of the
	// SMEM k_seed_fast already stored for the call) and the length d of its longest forward match; L-first: that match has K
	// or more bases and is walked first, its interval in defer_lx[]; walk = 0: for k_seed.
	// NULL: k_seed runs in read mode and takes every read.
	uint4 *defer_q;
	uint32_t *defer_bits;
	uint4 *thread_lx;           // [n_threads] where k_seed_fast keeps such an interval until the task has its queue slot
	uint4 *defer_lx;            // walk tasks that start with L itself (bit 27 of .y): its packed interval (pack_entry; x2 == 0: not known)
	uint32_t *lit_q, *n_lit;    // the walk = 0 entries of defer_q (indices), listed for k_seed
	uint32_t defer_cap;
	uint32_t *n_defer;
	const uint32_t *n_defer_fast; // n_defer as it stood when k_seed_fast ended (copied on the stream): what the walk tasks are picked from
	const uint32_t *walk_order;   // [n_walk] the walk tasks among them (indices into defer_q), longest first (k_walk_count / _scan / _scatter)
	const uint32_t *n_walk;
	uint32_t *read_last_q;      // [n_reads] last deferred call of each read (chain head), ~0 if none
	uint64_t *x_off;            // [defer_cap] where the mems of a deferred call start in pool
	uint32_t *x_n;              // [defer_cap]
	cs_mem_t *thread_mems;      // [n_threads][mem_cap] per-thread mem list of the read in flight
	uint32_t mem_cap;
	uint4 *spill;               // [spill_cap][n_threads] interval-list entries beyond CS_LIST_SMEM
	uint32_t spill_cap;
	// outputs
	cs_mem_t *pool;             // unsorted mems, reads in completion order
	uint64_t pool_cap;
	unsigned long long *pool_used;
	uint64_t *read_pool_off;    // [n_reads] where the read's mems start in pool
	uint32_t *read_n_mems;      // [n_reads]
	cs_mem_t *r3_mems;          // third-pass seeds of read r at [off[r]/(k+1) + r ...]
	uint32_t *r3_n_mems;        // [n_reads]
	unsigned long long *counters; // [0] ext queries [1] ext calls (bucket path) [2] two-sector extends [3] occurrence-filter probes
	unsigned long long *req;    // executed memory requests: [0] k_seed_fast [1] k_seed_walk [2] k_seed [3] third-pass kernel
	int *error;                 // sticky CS_E_* code (atomicMin: the per-read code CS_E_READ_OVERFLOW wins over CS_E_OVERFLOW)
};

struct CollectArgs {
	uint32_t n_reads;
	cs_seed_opt_t opt;
	const cs_mem_t *pool;
	const uint64_t *read_pool_off;
	const uint32_t *read_n_mems;
	const uint32_t *off;        // read offsets (locates the third-pass seeds)
	const cs_mem_t *r3_mems;
	const uint32_t *r3_n_mems;  // NULL when the third pass is disabled (max_mem_intv == 0)
	const uint32_t *read_last_q; // deferred calls of each read (NULL: none): chain through defer_q[].w, results at x_off / x_n
	const uint4 *defer_q;
	const uint64_t *x_off;
	const uint32_t *x_n;
	cs_mem_t *stage;            // scratch with the layout of `mems`: a read's sources gathered before the sort
	const uint32_t *mem_off;    // exclusive scan of the per-read totals, n_reads + 1
	cs_mem_t *mems;             // sorted output
	uint64_t mems_cap;
	uint32_t *read_n_seeds;     // [n_reads]
	const uint32_t *seed_off;   // exclusive scan of read_n_seeds (for pass 2)
	uint64_t *seed_rows;        // [n_seeds] SA rows in emission order (pass 2), resolved in place by k_sa_resolve
	uint64_t seed_cap;
	unsigned long long *tot_seeds;  // 64-bit total of read_n_seeds (the u32 offsets scan could wrap silently)
	int *error;
};

#include "cs_chain.cuh"

__global__ void k_relayout(const uint32_t *src, uint64_t src_words, uint64_t seq_len, uint4 *dst, uint64_t n_buckets);
__global__ void k_unlayout(const uint4 *src, uint64_t seq_len, uint32_t *dst, uint64_t dst_words);
__global__ void k_resample_sa(DevIndex I, uint64_t *out, uint64_t n_out, uint32_t out_shift);
__global__ void k_probe_occ4(DevIndex I, uint32_t n, const uint64_t *k, uint64_t *cnt);
__global__ void k_probe_extend(DevIndex I, uint32_t n, const uint64_t *ik, const int32_t *is_back, uint64_t *ok);
__global__ void k_sa_resolve(DevIndex I, const uint32_t *n_ptr, uint64_t cap, uint64_t *rows_inout, unsigned long long *work,
                             unsigned long long *lf_steps);
__global__ void k_kt_build(DevIndex I, uint4 *kt, uint32_t d);
__global__ void k_text_from_index(DevIndex I, unsigned long long *W);
__global__ void k_text_lsb(const uint64_t *W, uint64_t n_words, uint64_t *out);
__global__ void k_isa_sample(DevIndex I, uint64_t *isa, uint32_t shift);
#define CS_HAVE_REP 1
__global__ void k_rep_build(DevIndex I, uint8_t *rep);
__global__ void k_pt_count(const uint64_t *W, uint64_t n, uint32_t K, uint32_t *pt);
__global__ void k_pack_reads(const uint8_t *bases, const uint32_t *off, uint32_t n_reads, uint64_t *packed, uint32_t *nmask);
__global__ void k_unpack_reads(const uint64_t *packed, const uint32_t *nmask, const uint32_t *off, uint32_t n_reads, uint8_t *bases, uint32_t off_bias);
__global__ void k_seed(DevIndex I, SeedArgs a);
__global__ void k_seed_long(DevIndex I, SeedArgs a);
__global__ void k_seed_fast(DevIndex I, SeedArgs a);
__global__ void k_seed_walk(DevIndex I, SeedArgs a);
__global__ void k_walk_count(DevIndex I, SeedArgs a, uint32_t *hist);
__global__ void k_walk_scan(uint32_t *hist, uint32_t *cursor, uint32_t *n_walk);
__global__ void k_walk_scatter(DevIndex I, SeedArgs a, uint32_t *cursor, uint32_t *order);
__global__ void k_seed_r3_fast(DevIndex I, SeedArgs a);
__global__ void k_seed_r3(DevIndex I, SeedArgs a);
__global__ void k_mem_counts(const uint32_t *n12, const uint32_t *n3, const uint32_t *read_last_q, const uint4 *defer_q, const uint32_t *x_n,
                             uint32_t n_reads, uint32_t *out, unsigned long long *tot12, unsigned long long *tot3);
__global__ void k_collect_sort(CollectArgs a);
__global__ void k_collect_rows(CollectArgs a);
__global__ void k_compact_results(const uint32_t *n_mems_ptr, uint64_t mems_cap, const cs_mem_t *mems, cs_cmem_t *cmems,
                                  const uint32_t *n_seeds_ptr, uint64_t seeds_cap, const uint64_t *rbeg, uint32_t *lo, uint8_t *hi);
__global__ void k_gather_probe(const uint4 *table, uint64_t n_granules, uint32_t granule16, uint64_t n_loads, uint64_t seed, unsigned long long *sink);
__global__ void k_gather_probe4(const uint4 *table, uint64_t n_granules, uint32_t granule16, uint64_t n_loads, uint64_t seed, unsigned long long *sink);
__global__ void k_fill(uint4 *p, uint64_t n, uint32_t v);
