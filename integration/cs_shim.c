/* Host shim between the reference's unchanged C pipeline (mapping/bwamem.c, mapping/fastmap.c) and the C-ABI of
 * compseed_b200 (include/compseed_b200.h).  Compiled against the reference's own headers.
 *
 * Flow (north_star: "kthread workers swapped for a pinned-buffer, multi-stream batch pipeline"):
 *   step 0 of kt_pipeline (fastmap.c:76-99, the reader)   csgpu_prefetch_batch: the batch just read is converted to nt4 into a
 *                                                         page-locked buffer and handed to cs_multi_submit -- which returns at
 *                                                         once; the GPUs seed batch i+1 while the host chains / extends batch i
 *   step 1, mem_process_seqs before kt_for (bwamem.c:1343) csgpu_seed_batch: waits for that set (or seeds it now if nobody
 *                                                         prefetched it, e.g. the SMARTPE split of fastmap.c:104-121)
 *   worker1 -> mem_align1_core -> mem_chain               csgpu_set_read / csgpu_fill_mems / csgpu_next_rbeg read the results of
 *                                                         the read in hand where the DMA left them (compact wire format,
 *                                                         expanded per read by the kt_for workers, in parallel)
 * All visible GPUs are used (CSGPU_DEVICES=n limits them): the index is uploaded once and replicated device-to-device, each
 * GPU seeds one contiguous block of the batch, nothing is exchanged between them.
 * Errors follow the reference's convention: fatal (bwalib/utils.c:92-124). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "cs_shim.h"
#include "FM_index/bntseq.h"
#include "compseed_b200.h"

#define N_SETS 2            /* kt_pipeline runs two batches at a time (fastmap.c:379, kthread.c:95-107) */

typedef struct {
	const bseq1_t *seqs; int n;       /* which batch this set holds (identity of a prefetched batch) */
	int submitted;
	uint8_t *bases; uint64_t bases_cap; /* page-locked */
	uint64_t *off; uint64_t off_cap;
	uint32_t max_len;
} set_t;

static cs_index_t *g_idx[CS_MULTI_MAX_DEV];
static int g_ndev;
static cs_multi_t *g_multi;
static uint32_t g_multi_len;
static set_t g_set[N_SETS];
static int g_next_set;
static cs_multi_result_t g_res;     /* the batch the kt_for workers are reading */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;   /* step 0 (prefetch) and step 1 (seed) of different batches run concurrently */
static int g_chain = -1;            /* CSGPU_CHAIN=1: the GPUs also chain; the host takes chains instead of mems + seeds */
static __thread const cs_chain_t *t_ch; static __thread uint32_t t_nch; static __thread const cs_block_t *t_blk; static __thread uint64_t t_cs0;
static __thread const cs_cmem_t *t_cm; static __thread uint32_t t_nm;
static __thread const uint32_t *t_lo; static __thread const uint8_t *t_hi; static __thread uint64_t t_cursor;

static void fatal(const char *what)
{
	fprintf(stderr, "[csgpu] %s: %s\n", what, cs_last_error());
	exit(EXIT_FAILURE);
}

static void ensure_index(const bwt_t *bwt)
{ /* the index the host already loaded (bwa_idx_load, bwa.c:288) -> HBM of the first GPU, then device-to-device to the others */
	cs_bwt_view_t v;
	const char *dense = getenv("CSGPU_SA_INTV"), *devs = getenv("CSGPU_DEVICES");
	int k, want;
	if (g_ndev) return;
	want = cs_device_count();
	if (want < 1) { fprintf(stderr, "[csgpu] no CUDA device: this build has no CPU seeding path\n"); exit(EXIT_FAILURE); }
	if (devs && strcmp(devs, "all") != 0 && atoi(devs) > 0 && atoi(devs) < want) want = atoi(devs);
	if (want > CS_MULTI_MAX_DEV) want = CS_MULTI_MAX_DEV;
	memset(&v, 0, sizeof v);
	v.primary = bwt->primary; memcpy(v.L2, bwt->L2, sizeof v.L2); v.seq_len = bwt->seq_len;
	v.bwt_size = bwt->bwt_size; v.bwt = bwt->bwt; v.sa_intv = bwt->sa_intv; v.n_sa = bwt->n_sa; v.sa = bwt->sa;
	g_idx[0] = cs_index_upload(&v, 0, dense ? atoi(dense) : 1);
	if (!g_idx[0]) fatal("cs_index_upload");
	for (k = 1; k < want; ++k)
		if (!(g_idx[k] = cs_index_replicate(g_idx[0], k))) fatal("cs_index_replicate");
	g_ndev = want;
}

/* (re)creates the pipeline for reads up to max_len bases.  Only step 1 may re-create (no kt_for worker is reading results
 * then); a set step 0 prefetched for the next batch is dropped and will be seeded again when its own step 1 comes. */
static const bntseq_t *g_bns; static mem_opt_t g_chain_opt;

static void apply_chaining(void)
{ /* contig table (the fields bns_intv2rid reads) and the chaining scalars of mem_opt_t */
	cs_bns_view_t v; cs_chain_opt_t co;
	int64_t *off; uint8_t *alt; int i;
	if (g_chain <= 0 || !g_bns) return;
	off = (int64_t*)malloc(g_bns->n_seqs * 8); alt = (uint8_t*)malloc(g_bns->n_seqs);
	for (i = 0; i < g_bns->n_seqs; ++i) { off[i] = g_bns->anns[i].offset; alt[i] = g_bns->anns[i].is_alt != 0; }
	v.l_pac = g_bns->l_pac; v.n_seqs = g_bns->n_seqs; v.offset = off; v.is_alt = alt;
	co.w = g_chain_opt.w; co.max_chain_gap = g_chain_opt.max_chain_gap; co.min_chain_weight = g_chain_opt.min_chain_weight;
	co.max_chain_extend = g_chain_opt.max_chain_extend; co.mask_level = g_chain_opt.mask_level; co.drop_ratio = g_chain_opt.drop_ratio;
	if (cs_multi_set_chaining(g_multi, &v, &co) != CS_OK) fatal("cs_multi_set_chaining");
	free(off); free(alt);
}

static int ensure_multi(uint32_t max_len, int may_recreate)
{
	const char *b = getenv("CSGPU_BATCH");
	int s;
	if (g_multi && max_len <= g_multi_len) return 1;
	if (g_multi) {
		if (!may_recreate) return 0;
		for (s = 0; s < N_SETS; ++s)
			if (g_set[s].submitted) { cs_multi_result_t r; if (cs_multi_wait(g_multi, s, &r) != CS_OK) fatal("cs_multi_wait"); g_set[s].submitted = 0; }
		cs_multi_free(g_multi);
	}
	g_multi_len = max_len < 256 ? 256 : max_len;
	g_multi = cs_multi_create(g_idx, g_ndev, b && atoi(b) > 0 ? (uint32_t)atoi(b) : (1u << 18), g_multi_len, 3, 0, 0, NULL);
	if (!g_multi) fatal("cs_multi_create");
	apply_chaining();
	return 1;
}

/* nt4 conversion (bwamem.c:1176-1177 does the same later, in place) into the set's page-locked buffer, then submit */
static int submit_set(int s, const mem_opt_t *opt, int n, const bseq1_t *seqs, int may_recreate)
{
	set_t *t = &g_set[s];
	cs_seed_opt_t so;
	uint64_t total = 0;
	int i;
	t->max_len = 1;
	for (i = 0; i < n; ++i) { total += seqs[i].l_seq; if ((uint32_t)seqs[i].l_seq > t->max_len) t->max_len = seqs[i].l_seq; }
	if (total + 64 > t->bases_cap) {
		if (t->bases) { cs_host_unregister(t->bases); free(t->bases); }
		t->bases_cap = (total + 64) * 5 / 4;
		t->bases = (uint8_t*)malloc(t->bases_cap);
		if (!t->bases || cs_host_register(t->bases, t->bases_cap) != CS_OK) fatal("page-locking the read buffer");
	}
	if ((uint64_t)n + 1 > t->off_cap) { t->off_cap = ((uint64_t)n + 1) * 5 / 4; t->off = (uint64_t*)realloc(t->off, t->off_cap * 8); }
	t->off[0] = 0;
	for (i = 0; i < n; ++i) {
		const uint8_t *q = (const uint8_t*)seqs[i].seq;
		uint8_t *d = t->bases + t->off[i];
		int j;
		for (j = 0; j < seqs[i].l_seq; ++j) d[j] = q[j] < 4 ? q[j] : nst_nt4_table[q[j]];
		t->off[i + 1] = t->off[i] + seqs[i].l_seq;
	}
	so.min_seed_len = opt->min_seed_len;
	so.split_len = (int)(opt->min_seed_len * opt->split_factor + .499); /* bwamem.c:223 */
	so.split_width = opt->split_width; so.max_mem_intv = (int32_t)opt->max_mem_intv; so.max_occ = opt->max_occ;
	if (!ensure_multi(t->max_len, may_recreate)) return 0;
	if (cs_multi_submit(g_multi, s, (uint64_t)n, t->bases, t->off, &so) != CS_OK) fatal("cs_multi_submit");
	t->seqs = seqs; t->n = n; t->submitted = 1;
	return 1;
}

static void note_chain_setup(const mem_opt_t *opt, const bntseq_t *bns)
{
	if (g_chain < 0) g_chain = getenv("CSGPU_CHAIN") != 0 && atoi(getenv("CSGPU_CHAIN")) != 0;
	g_bns = bns; g_chain_opt = *opt;
}

int csgpu_chaining(void) { return g_chain > 0; }

void csgpu_prefetch_batch(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, int n, const bseq1_t *seqs)
{
	static int off = -1;
	int s;
	if (off < 0) off = getenv("CSGPU_NO_PREFETCH") != 0;
	if (n <= 0 || off) return;
	pthread_mutex_lock(&g_mu);
	note_chain_setup(opt, bns);
	ensure_index(bwt);
	s = g_next_set;
	if (!g_set[s].submitted && submit_set(s, opt, n, seqs, 0)) g_next_set = (g_next_set + 1) % N_SETS;
	/* (else: both sets busy, or a read longer than the pipeline was built for: step 1 seeds this batch itself) */
	pthread_mutex_unlock(&g_mu);
}

void csgpu_seed_batch(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, int n, const bseq1_t *seqs)
{
	int s, found = -1;
	if (n <= 0) return;
	pthread_mutex_lock(&g_mu);
	note_chain_setup(opt, bns);
	ensure_index(bwt);
	for (s = 0; s < N_SETS; ++s)
		if (g_set[s].submitted == 1 && g_set[s].seqs == seqs && g_set[s].n == n) found = s;
	if (found < 0) { /* not prefetched: seed it now, in whichever set is free */
		for (s = 0; s < N_SETS && found < 0; ++s) if (!g_set[s].submitted) found = s;
		if (found < 0) { fprintf(stderr, "[csgpu] no free read set\n"); exit(EXIT_FAILURE); }
		submit_set(found, opt, n, seqs, 1);
		if (!g_set[found].submitted) { /* the pipeline was re-created and this very set dropped meanwhile: cannot happen for the set just submitted */
			fprintf(stderr, "[csgpu] internal: set lost\n"); exit(EXIT_FAILURE);
		}
	}
	pthread_mutex_unlock(&g_mu);
	if (cs_multi_wait(g_multi, found, &g_res) != CS_OK) fatal("cs_multi_wait");
	pthread_mutex_lock(&g_mu);
	g_set[found].submitted = 0;                 /* (its result arrays stay valid until the set is submitted again) */
	pthread_mutex_unlock(&g_mu);
}

void csgpu_set_read(int i)
{
	uint32_t ns;
	if (g_chain > 0) cs_multi_read_chains(&g_res, (uint64_t)i, &t_ch, &t_nch, &t_blk, &t_cs0);
	else cs_multi_read(&g_res, (uint64_t)i, &t_cm, &t_nm, &t_lo, &t_hi, &t_cursor, &ns);
}

/* chains of the read in hand (CSGPU_CHAIN=1), for the glue in integration/bwamem_chain_glue.c */
int csgpu_n_chains(void) { return (int)t_nch; }
void csgpu_chain(int c, int *rid, int *w, int *kept, int *is_alt, int *n, int *l_rep, uint64_t *first_seed)
{
	const cs_chain_t *p = t_ch + c;
	uint64_t s = t_cs0; int k;
	for (k = 0; k < c; ++k) s += t_ch[k].n;
	*rid = p->rid; *w = (int)(p->w_kept & 0x1fffffffu); *kept = (int)((p->w_kept >> 29) & 3); *is_alt = (int)(p->w_kept >> 31);
	*n = (int)p->n; *l_rep = (int)p->l_rep; *first_seed = s;
}
void csgpu_chain_seed(uint64_t s, int64_t *rbeg, int *qbeg, int *len)
{
	*rbeg = cs_crbeg(t_blk->rbeg_lo, t_blk->rbeg_hi, s); *qbeg = t_blk->qbeg[s]; *len = t_blk->len[s];
}

void csgpu_fill_mems(bwtintv_v *mem)
{
	uint32_t k;
	if (t_nm > mem->m) { mem->m = t_nm; mem->a = (bwtintv_t*)realloc(mem->a, mem->m * sizeof(bwtintv_t)); }
	for (k = 0; k < t_nm; ++k) cs_cmem_unpack(t_cm + k, (cs_mem_t*)(mem->a + k)); /* cs_mem_t == bwtintv_t, bwt.h:62-64 */
	mem->n = t_nm;
}

int64_t csgpu_next_rbeg(void) { return cs_crbeg(t_lo, t_hi, t_cursor++); }

void csgpu_destroy(void)
{
	int k;
	if (g_multi) cs_multi_free(g_multi);
	for (k = 0; k < g_ndev; ++k) cs_index_free(g_idx[k]);
	for (k = 0; k < N_SETS; ++k) {
		if (g_set[k].bases) { cs_host_unregister(g_set[k].bases); free(g_set[k].bases); }
		free(g_set[k].off);
	}
	memset(g_set, 0, sizeof g_set);
	g_multi = 0; g_ndev = 0;
}
