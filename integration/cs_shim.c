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

void csgpu_prefetch_batch(const mem_opt_t *opt, const bwt_t *bwt, int n, const bseq1_t *seqs)
{
	static int off = -1;
	int s;
	if (off < 0) off = getenv("CSGPU_NO_PREFETCH") != 0;
	if (n <= 0 || off) return;
	pthread_mutex_lock(&g_mu);
	ensure_index(bwt);
	s = g_next_set;
	if (!g_set[s].submitted && submit_set(s, opt, n, seqs, 0)) g_next_set = (g_next_set + 1) % N_SETS;
	/* (else: both sets busy, or a read longer than the pipeline was built for: step 1 seeds this batch itself) */
	pthread_mutex_unlock(&g_mu);
}

void csgpu_seed_batch(const mem_opt_t *opt, const bwt_t *bwt, int n, const bseq1_t *seqs)
{
	int s, found = -1;
	if (n <= 0) return;
	pthread_mutex_lock(&g_mu);
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
	cs_multi_read(&g_res, (uint64_t)i, &t_cm, &t_nm, &t_lo, &t_hi, &t_cursor, &ns);
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
