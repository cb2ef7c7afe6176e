/* Host shim between the reference's unchanged C pipeline (mapping/bwamem.c) and the C-ABI of
 * compseed_b200 (include/compseed_b200.h).  Compiled against the reference's own headers.
 *
 * Batch flow: csgpu_seed_batch() converts the batch to nt4 (as bwamem.c:1176-1177 does later, in
 * place), pushes it through the slots of one context in chunks (chunk i+1 is submitted before
 * chunk i is waited on) and keeps the results; worker threads then read them by read index.
 * Errors follow the reference's convention: fatal (bwalib/utils.c:92-124). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cs_shim.h"
#include "FM_index/bntseq.h"
#include "compseed_b200.h"

#define CHUNK_READS (1u << 18)
#define N_SLOTS 2

static cs_index_t *g_idx;
static cs_ctx_t *g_ctx;
static uint32_t g_ctx_reads, g_ctx_len; static uint64_t g_ctx_bases, g_ctx_mems, g_ctx_seeds;
static uint32_t *g_mem_off, *g_seed_off;   /* n + 1, batch-global */
static cs_mem_t *g_mems; static int64_t *g_rbeg;
static uint64_t g_mems_cap, g_rbeg_cap; static int g_off_cap;
static __thread int t_read;
static __thread uint64_t t_cursor;

static void fatal(const char *what)
{
	fprintf(stderr, "[csgpu] %s: %s\n", what, cs_last_error());
	exit(EXIT_FAILURE);
}

static void ensure_ctx(uint32_t reads, uint64_t bases, uint32_t max_len, uint64_t mems, uint64_t seeds)
{
	if (g_ctx && reads <= g_ctx_reads && bases <= g_ctx_bases && max_len <= g_ctx_len && mems <= g_ctx_mems && seeds <= g_ctx_seeds) return;
	if (g_ctx) cs_ctx_free(g_ctx);
	if (reads > g_ctx_reads) g_ctx_reads = reads;
	if (bases > g_ctx_bases) g_ctx_bases = bases;
	if (max_len > g_ctx_len) g_ctx_len = max_len;
	if (mems > g_ctx_mems) g_ctx_mems = mems;
	if (seeds > g_ctx_seeds) g_ctx_seeds = seeds;
	g_ctx = cs_ctx_create(g_idx, g_ctx_reads, g_ctx_bases, g_ctx_len, g_ctx_mems, g_ctx_seeds, N_SLOTS);
	if (!g_ctx) fatal("cs_ctx_create");
}

void csgpu_seed_batch(const mem_opt_t *opt, const bwt_t *bwt, int n, const bseq1_t *seqs)
{
	cs_seed_opt_t so;
	uint64_t total = 0, m_used = 0, s_used = 0, need_m = 0, need_s = 0;
	static int packed_input = -1;   /* CSGPU_PACKED, read once */
	uint32_t max_len = 1, *off;
	uint8_t *bases;
	int i, c, n_chunks, next, done;
	if (n <= 0) return;
	if (packed_input < 0) packed_input = getenv("CSGPU_PACKED") != 0;
	if (!g_idx) { /* the index the host already loaded (bwa_idx_load, bwa.c:288) -> HBM, once */
		cs_bwt_view_t v;
		const char *dense = getenv("CSGPU_SA_INTV");
		memset(&v, 0, sizeof v);
		v.primary = bwt->primary; memcpy(v.L2, bwt->L2, sizeof v.L2); v.seq_len = bwt->seq_len;
		v.bwt_size = bwt->bwt_size; v.bwt = bwt->bwt; v.sa_intv = bwt->sa_intv; v.n_sa = bwt->n_sa; v.sa = bwt->sa;
		g_idx = cs_index_upload(&v, 0, dense ? atoi(dense) : 1);
		if (!g_idx) fatal("cs_index_upload");
	}
	so.min_seed_len = opt->min_seed_len;
	so.split_len = (int)(opt->min_seed_len * opt->split_factor + .499); /* bwamem.c:223 */
	so.split_width = opt->split_width; so.max_mem_intv = (int32_t)opt->max_mem_intv; so.max_occ = opt->max_occ;
	for (i = 0; i < n; ++i) { total += seqs[i].l_seq; if ((uint32_t)seqs[i].l_seq > max_len) max_len = seqs[i].l_seq; }
	bases = (uint8_t*)malloc(total + 1);
	off = (uint32_t*)malloc(((size_t)n + 1) * 4);
	off[0] = 0;
	for (i = 0; i < n; ++i) {
		const uint8_t *s = (const uint8_t*)seqs[i].seq;
		uint8_t *d = bases + off[i];
		int j;
		for (j = 0; j < seqs[i].l_seq; ++j) d[j] = s[j] < 4 ? s[j] : nst_nt4_table[s[j]]; /* bwamem.c:1176-1177 */
		off[i + 1] = off[i] + seqs[i].l_seq;
	}
	if (n + 1 > g_off_cap) {
		g_off_cap = n + 1;
		g_mem_off = (uint32_t*)realloc(g_mem_off, (size_t)g_off_cap * 4);
		g_seed_off = (uint32_t*)realloc(g_seed_off, (size_t)g_off_cap * 4);
	}
	g_mem_off[0] = g_seed_off[0] = 0;
	n_chunks = (n + CHUNK_READS - 1) / CHUNK_READS;
	for (;;) { /* retried with larger result buffers if a chunk overflows them */
		uint32_t per = n < (int)CHUNK_READS ? (uint32_t)n : CHUNK_READS;
		uint64_t cb = 0;
		int overflow = 0;
		for (c = 0; c < n_chunks; ++c) {
			int s = c * CHUNK_READS, e = s + CHUNK_READS < n ? s + CHUNK_READS : n;
			if (off[e] - off[s] > cb) cb = off[e] - off[s];
		}
		/* capacities follow the chunk size (16 mems / 32 seeds per read) until a batch has said what it needs */
		ensure_ctx(per, cb ? cb : 1, max_len, g_ctx_mems > (uint64_t)per * 16 ? g_ctx_mems : (uint64_t)per * 16,
		           g_ctx_seeds > (uint64_t)per * 32 ? g_ctx_seeds : (uint64_t)per * 32);
		m_used = s_used = 0;
		for (next = 0, done = 0; done < n_chunks && !overflow; ) {
			while (next < n_chunks && next - done < N_SLOTS) { /* keep every slot busy */
				int s = next * CHUNK_READS, e = s + CHUNK_READS < n ? s + CHUNK_READS : n, r;
				uint32_t *lo = (uint32_t*)malloc(((size_t)(e - s) + 1) * 4);
				for (r = s; r <= e; ++r) lo[r - s] = off[r] - off[s];
				if (!packed_input) {
					if (cs_seed_batch_submit(g_ctx, next % N_SLOTS, (uint32_t)(e - s), bases + off[s], lo, &so) != CS_OK) fatal("cs_seed_batch_submit");
				} else { /* CSGPU_PACKED=1: send the chunk 2-bit packed (57 instead of 150 bytes per 150-bp read cross the link) */
					const uint64_t nw = cs_packed_words((uint32_t)(e - s), lo);
					uint64_t *pk = (uint64_t*)malloc((nw ? nw : 1) * 8);
					uint32_t *nm = (uint32_t*)malloc((nw ? nw : 1) * 4);
					if (cs_pack_reads_host((uint32_t)(e - s), bases + off[s], lo, pk, nm, opt->n_threads) != CS_OK) fatal("cs_pack_reads_host");
					if (cs_seed_batch_submit_packed(g_ctx, next % N_SLOTS, (uint32_t)(e - s), pk, nm, lo, &so) != CS_OK) fatal("cs_seed_batch_submit_packed");
					free(pk); free(nm);                                   /* (the library staged them in its own pinned buffers) */
				}
				free(lo);
				++next;
			}
			{
				cs_result_t res;
				int s = done * CHUNK_READS, r, rc = cs_seed_batch_wait(g_ctx, done % N_SLOTS, &res);
				if (rc == CS_E_OVERFLOW) { overflow = 1; cs_ctx_need(g_ctx, done % N_SLOTS, &need_m, &need_s); break; }   /* (CS_E_READ_OVERFLOW is fatal: no buffer size fixes it) */
				if (rc != CS_OK) fatal("cs_seed_batch_wait");
				if (m_used + res.n_mems > g_mems_cap) { g_mems_cap = (m_used + res.n_mems) * 2; g_mems = (cs_mem_t*)realloc(g_mems, g_mems_cap * sizeof(cs_mem_t)); }
				if (s_used + res.n_seeds > g_rbeg_cap) { g_rbeg_cap = (s_used + res.n_seeds) * 2; g_rbeg = (int64_t*)realloc(g_rbeg, g_rbeg_cap * 8); }
				memcpy(g_mems + m_used, res.mems, res.n_mems * sizeof(cs_mem_t));
				memcpy(g_rbeg + s_used, res.rbeg, res.n_seeds * 8);
				for (r = 1; r <= (int)res.n_reads; ++r) {
					g_mem_off[s + r] = (uint32_t)(m_used + res.mem_off[r]);
					g_seed_off[s + r] = (uint32_t)(s_used + res.seed_off[r]);
				}
				m_used += res.n_mems; s_used += res.n_seeds;
				++done;
			}
		}
		if (!overflow) break;
		/* drain the slots still in flight, then grow and redo the batch */
		for (c = done + 1; c < next; ++c) { cs_result_t res; cs_seed_batch_wait(g_ctx, c % N_SLOTS, &res); }
		if (need_m >= (1ull << 32) || need_s >= (1ull << 32)) { fprintf(stderr, "[csgpu] a chunk needs more than 2^32 mems or seeds\n"); exit(EXIT_FAILURE); }
		if (need_m > g_ctx_mems) g_ctx_mems = need_m;      /* what the library said this chunk needs: one retry is enough */
		if (need_s > g_ctx_seeds) g_ctx_seeds = need_s;
		cs_ctx_free(g_ctx); g_ctx = 0;
	}
	free(bases); free(off);
}

void csgpu_set_read(int i) { t_read = i; t_cursor = g_seed_off[i]; }

void csgpu_fill_mems(bwtintv_v *mem)
{
	size_t n = g_mem_off[t_read + 1] - g_mem_off[t_read];
	if (n > mem->m) { mem->m = n; mem->a = (bwtintv_t*)realloc(mem->a, mem->m * sizeof(bwtintv_t)); }
	memcpy(mem->a, g_mems + g_mem_off[t_read], n * sizeof(bwtintv_t)); /* cs_mem_t == bwtintv_t, bwt.h:62-64 */
	mem->n = n;
}

int64_t csgpu_next_rbeg(void) { return g_rbeg[t_cursor++]; }

void csgpu_destroy(void)
{
	if (g_ctx) cs_ctx_free(g_ctx);
	if (g_idx) cs_index_free(g_idx);
	g_ctx = 0; g_idx = 0;
	free(g_mem_off); free(g_seed_off); free(g_mems); free(g_rbeg);
}
