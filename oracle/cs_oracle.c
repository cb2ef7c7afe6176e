/* TEST INFRASTRUCTURE ONLY -- see cs_oracle.h for the usage rules and the parity status (PINNED).
 *
 * CPU restatement of the reference's SMEM seeding path.  Written from the behaviour described in
 * SURVEY.md section 8a, each function citing the reference lines it restates.  It is deliberately
 * plain scalar C: it is the checker, not a product.
 */
#define _GNU_SOURCE
#include "cs_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * Index construction (restates the *result* of bwa_idx_build, FM_index/index_main.c:257-325):
 *   T = fwd + revcomp(fwd)                      (bns_fasta2bntseq for_only=0, bntseq.c:314-318)
 *   BWT of T$ with '$' smallest, '$' row dropped (bwt_pac2bwt, index_main.c:60-131)
 *   Occ checkpoints interleaved every 128 rows  (bwt_bwtupdate_core, index_main.c:152-174)
 *   SA sampled every sa_intv rows, sa[0] = -1   (bwt_cal_sa, bwt.c:62-84)
 * The suffix sort is a prefix-doubling sort of our own (the reference uses SA-IS / BWT-SW; any
 * correct suffix sort yields the same BWT).
 * ------------------------------------------------------------------------------------------ */

typedef struct { int32_t key; uint32_t sa; } pair_t;
static int pair_cmp(const void *a, const void *b)
{
	int32_t x = ((const pair_t*)a)->key, y = ((const pair_t*)b)->key;
	return (x > y) - (x < y);
}

typedef struct { uint32_t s, e; } grp_t;

static uint32_t *suffix_sort(const uint8_t *T, uint32_t n)
{
	enum { K = 10 };
	const uint32_t NB = 9765625u; /* 5^10 */
	const uint32_t TOP = 1953125u; /* 5^9 */
	uint32_t *SA = (uint32_t*)malloc((size_t)n * 4);
	int32_t *rank = (int32_t*)malloc(((size_t)n + 1) * 4);
	uint32_t *key = (uint32_t*)malloc((size_t)n * 4);
	uint32_t *cnt = (uint32_t*)calloc((size_t)NB + 1, 4);
	uint32_t i, k = 0;
	/* base-5 key of the first K symbols, '$'/padding = digit 0, so truncated suffixes get unique keys */
	for (i = n; i-- > 0; ) { k = k / 5 + (uint32_t)(T[i] + 1) * TOP; key[i] = k; }
	for (i = 0; i < n; ++i) ++cnt[key[i] + 1];
	for (i = 0; i < NB; ++i) cnt[i + 1] += cnt[i];
	for (i = 0; i < n; ++i) rank[i] = (int32_t)cnt[key[i]]; /* rank = start of the group */
	rank[n] = -1;
	{
		uint32_t *pos = (uint32_t*)malloc((size_t)NB * 4);
		memcpy(pos, cnt, (size_t)NB * 4);
		for (i = 0; i < n; ++i) SA[pos[key[i]]++] = i;
		free(pos);
	}
	/* unresolved groups */
	size_t ng = 0, mg = 1024;
	grp_t *g = (grp_t*)malloc(mg * sizeof(grp_t));
	for (i = 0; i < NB; ++i) {
		if (cnt[i + 1] - cnt[i] > 1) {
			if (ng == mg) { mg <<= 1; g = (grp_t*)realloc(g, mg * sizeof(grp_t)); }
			g[ng].s = cnt[i]; g[ng].e = cnt[i + 1]; ++ng;
		}
	}
	free(cnt); free(key);
	uint32_t h = K;
	pair_t *buf = 0; size_t mbuf = 0;
	int32_t *newrank = (int32_t*)malloc((size_t)n * 4); /* indexed by SA position */
	while (ng) {
		size_t ng2 = 0, mg2 = ng + 16, gi;
		grp_t *g2 = (grp_t*)malloc(mg2 * sizeof(grp_t));
		for (gi = 0; gi < ng; ++gi) {
			uint32_t s = g[gi].s, e = g[gi].e, m = e - s, p;
			if (m > mbuf) { mbuf = m * 2; buf = (pair_t*)realloc(buf, mbuf * sizeof(pair_t)); }
			for (p = 0; p < m; ++p) { buf[p].sa = SA[s + p]; buf[p].key = rank[(size_t)SA[s + p] + h]; }
			qsort(buf, m, sizeof(pair_t), pair_cmp);
			uint32_t gs = 0;
			for (p = 0; p < m; ++p) {
				SA[s + p] = buf[p].sa;
				if (p > 0 && buf[p].key != buf[p - 1].key) {
					if (p - gs > 1) {
						if (ng2 == mg2) { mg2 <<= 1; g2 = (grp_t*)realloc(g2, mg2 * sizeof(grp_t)); }
						g2[ng2].s = s + gs; g2[ng2].e = s + p; ++ng2;
					}
					gs = p;
				}
				newrank[s + p] = (int32_t)(s + gs);
			}
			if (m - gs > 1) {
				if (ng2 == mg2) { mg2 <<= 1; g2 = (grp_t*)realloc(g2, mg2 * sizeof(grp_t)); }
				g2[ng2].s = s + gs; g2[ng2].e = s + m; ++ng2;
			}
		}
		/* commit ranks only after the whole pass (all comparisons of a pass use the same ranks) */
		for (gi = 0; gi < ng; ++gi) {
			uint32_t p;
			for (p = g[gi].s; p < g[gi].e; ++p) rank[SA[p]] = newrank[p];
		}
		free(g); g = g2; ng = ng2;
		h <<= 1;
	}
	free(g); free(buf); free(newrank); free(rank);
	return SA;
}

cso_index_t *cso_index_build(const uint8_t *fwd, uint64_t l_pac, int sa_intv)
{
	uint64_t n = 2 * l_pac, i;
	if (n == 0 || n >= 0x7fffff00ull || sa_intv < 1 || (sa_intv & (sa_intv - 1))) return 0;
	uint8_t *T = (uint8_t*)malloc(n);
	for (i = 0; i < l_pac; ++i) { T[i] = fwd[i] & 3; T[n - 1 - i] = 3 - (fwd[i] & 3); }
	uint32_t *SA = suffix_sort(T, (uint32_t)n);
	cso_index_t *idx = (cso_index_t*)calloc(1, sizeof(cso_index_t));
	idx->seq_len = n;
	for (i = 0; i < n; ++i) ++idx->L2[T[i] + 1];
	for (i = 1; i <= 4; ++i) idx->L2[i] += idx->L2[i - 1];
	/* BWT string without the '$' row; full-matrix row r >= 1 is SA[r-1], row 0 is the '$' suffix */
	uint8_t *B = (uint8_t*)malloc(n);
	uint64_t j = 0;
	B[j++] = T[n - 1];
	for (i = 0; i < n; ++i) {
		if (SA[i] == 0) idx->primary = i + 1;
		else B[j++] = T[SA[i] - 1];
	}
	/* interleaved layout, index_main.c:152-174 */
	uint64_t n_occ = (n + 127) / 128 + 1;
	idx->bwt_size = ((n + 15) >> 4) + n_occ * 8;
	idx->bwt = (uint32_t*)calloc(idx->bwt_size, 4);
	uint64_t c[4] = {0, 0, 0, 0}, k = 0;
	for (i = 0; i < n; ++i) {
		if ((i & 127) == 0) { memcpy(idx->bwt + k, c, 32); k += 8; }
		if ((i & 15) == 0) ++k;
		idx->bwt[k - 1] |= (uint32_t)B[i] << ((15 - (i & 15)) << 1);
		++c[B[i]];
	}
	memcpy(idx->bwt + k, c, 32);
	if (k + 8 != idx->bwt_size) { fprintf(stderr, "[cso_index_build] inconsistent bwt_size\n"); abort(); }
	/* sampled SA, bwt.c:62-84 */
	idx->sa_intv = sa_intv;
	idx->n_sa = (n + sa_intv) / sa_intv;
	idx->sa = (uint64_t*)calloc(idx->n_sa, 8);
	idx->sa[0] = (uint64_t)-1;
	for (i = 1; i <= n; ++i)
		if (i % sa_intv == 0) idx->sa[i / sa_intv] = SA[i - 1];
	free(B); free(SA); free(T);
	return idx;
}

/* On-disk format: SURVEY.md Appendix B; bwt.c:385-407 (dump), bwt.c:421-462 (restore). */
cso_index_t *cso_index_load(const char *prefix)
{
	char fn[4096];
	FILE *fp;
	cso_index_t *idx = (cso_index_t*)calloc(1, sizeof(cso_index_t));
	snprintf(fn, sizeof fn, "%s.bwt", prefix);
	if ((fp = fopen(fn, "rb")) == 0) { free(idx); return 0; }
	fseek(fp, 0, SEEK_END);
	idx->bwt_size = ((uint64_t)ftell(fp) - 40) >> 2;
	fseek(fp, 0, SEEK_SET);
	idx->bwt = (uint32_t*)calloc(idx->bwt_size, 4);
	if (fread(&idx->primary, 8, 1, fp) != 1 || fread(idx->L2 + 1, 8, 4, fp) != 4 ||
	    fread(idx->bwt, 4, idx->bwt_size, fp) != idx->bwt_size) { fclose(fp); cso_index_free(idx); return 0; }
	fclose(fp);
	idx->seq_len = idx->L2[4];
	snprintf(fn, sizeof fn, "%s.sa", prefix);
	if ((fp = fopen(fn, "rb")) == 0) { cso_index_free(idx); return 0; }
	uint64_t hdr[7];
	if (fread(hdr, 8, 7, fp) != 7 || hdr[0] != idx->primary || hdr[6] != idx->seq_len) { fclose(fp); cso_index_free(idx); return 0; }
	idx->sa_intv = (int32_t)hdr[5];
	idx->n_sa = (idx->seq_len + idx->sa_intv) / idx->sa_intv;
	idx->sa = (uint64_t*)calloc(idx->n_sa, 8);
	idx->sa[0] = (uint64_t)-1;
	if (fread(idx->sa + 1, 8, idx->n_sa - 1, fp) != idx->n_sa - 1) { fclose(fp); cso_index_free(idx); return 0; }
	fclose(fp);
	return idx;
}

int cso_index_dump(const cso_index_t *idx, const char *prefix)
{
	char fn[4096];
	FILE *fp;
	uint64_t v;
	snprintf(fn, sizeof fn, "%s.bwt", prefix);
	if ((fp = fopen(fn, "wb")) == 0) return -1;
	fwrite(&idx->primary, 8, 1, fp); fwrite(idx->L2 + 1, 8, 4, fp); fwrite(idx->bwt, 4, idx->bwt_size, fp);
	fclose(fp);
	snprintf(fn, sizeof fn, "%s.sa", prefix);
	if ((fp = fopen(fn, "wb")) == 0) return -1;
	fwrite(&idx->primary, 8, 1, fp); fwrite(idx->L2 + 1, 8, 4, fp);
	v = (uint64_t)idx->sa_intv; fwrite(&v, 8, 1, fp);
	fwrite(&idx->seq_len, 8, 1, fp);
	fwrite(idx->sa + 1, 8, idx->n_sa - 1, fp);
	fclose(fp);
	return 0;
}

void cso_index_free(cso_index_t *idx)
{
	if (!idx) return;
	free(idx->bwt); free(idx->sa); free(idx);
}

/* ------------------------------------------------------------------------------------------
 * FM-index queries
 * ------------------------------------------------------------------------------------------ */

/* number of occurrences of each base among the top `nb` (1..16) bases of word w (MSB first) */
static inline void count16(uint32_t w, int nb, uint64_t cnt[4])
{
	int i;
	for (i = 0; i < nb; ++i) ++cnt[w >> ((15 - i) << 1) & 3];
}

/* bwt.c:169-186: cnt[c] = #c in BWT rows [0..k] inclusive; k == -1 -> zeros */
void cso_occ4(const cso_index_t *idx, uint64_t k, uint64_t cnt[4])
{
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= idx->primary); /* '$' is not stored */
	const uint32_t *p = idx->bwt + ((k >> 7) << 4);
	memcpy(cnt, p, 32);
	p += 8;
	int r = (int)(k & 127), w;
	for (w = 0; w < (r >> 4); ++w) count16(p[w], 16, cnt);
	count16(p[r >> 4], (r & 15) + 1, cnt);
}

static inline uint64_t occ1(const cso_index_t *idx, uint64_t k, int c) /* bwt.c:107-129 */
{
	uint64_t cnt[4];
	if (k == idx->seq_len) return idx->L2[c + 1] - idx->L2[c];
	if (k == (uint64_t)-1) return 0;
	cso_occ4(idx, k, cnt);
	return cnt[c];
}

/* bwt.c:262-275 (bwt_2occ4 bwt.c:189-220 yields the same numbers as two bwt_occ4 calls) */
void cso_extend(const cso_index_t *idx, const uint64_t ik[3], uint64_t ok[4][3], int is_back)
{
	uint64_t tk[4], tl[4];
	int i, a = !is_back, b = is_back;
	cso_occ4(idx, ik[a] - 1, tk);
	cso_occ4(idx, ik[a] - 1 + ik[2], tl);
	for (i = 0; i < 4; ++i) {
		ok[i][a] = idx->L2[i] + 1 + tk[i];
		ok[i][2] = tl[i] - tk[i];
	}
	ok[3][b] = ik[b] + (ik[a] <= idx->primary && ik[a] + ik[2] - 1 >= idx->primary);
	ok[2][b] = ok[3][b] + ok[3][2];
	ok[1][b] = ok[2][b] + ok[2][2];
	ok[0][b] = ok[1][b] + ok[1][2];
}

static inline int ext_two_buckets(const cso_index_t *idx, const uint64_t ik[3], int is_back)
{ /* the branch condition of bwt.c:192-194 */
	uint64_t k = ik[!is_back] - 1, l = k + ik[2];
	uint64_t _k = k - (k >= idx->primary), _l = l - (l >= idx->primary);
	return (_l >> 7 != _k >> 7) || k == (uint64_t)-1 || l == (uint64_t)-1;
}

/* bwt.c:53-59 + bwt.c:86-96 */
static uint64_t sa_lookup(const cso_index_t *idx, uint64_t k, int64_t *steps)
{
	uint64_t sa = 0, mask = (uint64_t)idx->sa_intv - 1;
	while (k & mask) {
		++sa;
		if (k == idx->primary) k = 0;
		else {
			uint64_t x = k - (k > idx->primary);
			int c = idx->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)] >> ((~x & 15) << 1) & 3;
			k = idx->L2[c] + occ1(idx, k, c);
		}
	}
	if (steps) *steps += (int64_t)sa;
	return sa + idx->sa[k / idx->sa_intv];
}
uint64_t cso_sa(const cso_index_t *idx, uint64_t k) { return sa_lookup(idx, k, 0); }

void cso_occ4_many(const cso_index_t *idx, int n, const uint64_t *k, uint64_t *cnt)
{ int i; for (i = 0; i < n; ++i) cso_occ4(idx, k[i], cnt + 4 * (size_t)i); }
void cso_extend_many(const cso_index_t *idx, int n, const uint64_t *ik, const int32_t *is_back, uint64_t *ok)
{ int i; for (i = 0; i < n; ++i) cso_extend(idx, ik + 3 * (size_t)i, (uint64_t(*)[3])(ok + 12 * (size_t)i), is_back[i]); }
void cso_sa_many(const cso_index_t *idx, int n, const uint64_t *k, uint64_t *out)
{ int i; for (i = 0; i < n; ++i) out[i] = cso_sa(idx, k[i]); }

/* ------------------------------------------------------------------------------------------
 * Seeding
 * ------------------------------------------------------------------------------------------ */

typedef struct { size_t n, m; cso_mem_t *a; } memv_t;
static inline void memv_push(memv_t *v, const cso_mem_t *x)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 16; v->a = (cso_mem_t*)realloc(v->a, v->m * sizeof(cso_mem_t)); }
	v->a[v->n++] = *x;
}
static void memv_reverse(memv_t *v)
{
	size_t i;
	for (i = 0; i < v->n >> 1; ++i) { cso_mem_t t = v->a[i]; v->a[i] = v->a[v->n - 1 - i]; v->a[v->n - 1 - i] = t; }
}

typedef struct {
	const cso_index_t *idx;
	memv_t prev, curr, mem1;
	int64_t cnt[CSO_N_CNT];
	int round;
} work_t;

static inline void set_intv(const cso_index_t *idx, int c, cso_mem_t *ik) /* bwt.h:82 */
{
	ik->x[0] = idx->L2[c] + 1; ik->x[2] = idx->L2[c + 1] - idx->L2[c]; ik->x[1] = idx->L2[3 - c] + 1; ik->info = 0;
}

/* optional profiling histogram (env CSO_HIST=1): extends by [round][direction][length of the extended string] */
static int64_t g_hist[3][2][64];
static int g_hist_on = -1;
void cso_hist_get(int64_t *out) { memcpy(out, g_hist, sizeof g_hist); }
void cso_hist_reset(void) { memset(g_hist, 0, sizeof g_hist); }

static inline void counted_extend_len(work_t *w, const cso_mem_t *ik, uint64_t ok[4][3], int is_back, int new_len)
{
	if (g_hist_on < 0) g_hist_on = getenv("CSO_HIST") != 0;
	if (g_hist_on) __atomic_fetch_add(&g_hist[w->round][is_back][new_len < 63 ? new_len : 63], 1, __ATOMIC_RELAXED);
	++w->cnt[CSO_CNT_EXT];
	++w->cnt[CSO_CNT_EXT_R1 + w->round];
	w->cnt[CSO_CNT_EXT2] += ext_two_buckets(w->idx, ik->x, is_back);
	cso_extend(w->idx, ik->x, ok, is_back);
}

static inline void counted_extend(work_t *w, const cso_mem_t *ik, uint64_t ok[4][3], int is_back)
{
	++w->cnt[CSO_CNT_EXT];
	++w->cnt[CSO_CNT_EXT_R1 + w->round];
	w->cnt[CSO_CNT_EXT2] += ext_two_buckets(w->idx, ik->x, is_back);
	cso_extend(w->idx, ik->x, ok, is_back);
}

/* bwt_smem1a with max_intv == 0 (bwt.c:289-351): SMEMs covering position x with >= min_intv hits.
 * Result in w->mem1 sorted by start; returns the end of the longest match from x (next pivot). */
static int smem1(work_t *w, int len, const uint8_t *q, int x, uint64_t min_intv)
{
	const cso_index_t *idx = w->idx;
	memv_t *prev = &w->prev, *curr = &w->curr, *swap, *mem = &w->mem1;
	cso_mem_t ik, t;
	uint64_t ok[4][3];
	int i, c, ret;
	size_t j;

	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	set_intv(idx, q[x], &ik);
	ik.info = (uint64_t)(x + 1);
	curr->n = 0;
	for (i = x + 1; i < len; ++i) { /* forward: record the interval every time its size changes */
		if (q[i] < 4) {
			c = 3 - q[i];
			counted_extend_len(w, &ik, ok, 0, i - x + 1);
			if (ok[c][2] != ik.x[2]) {
				memv_push(curr, &ik);
				if (ok[c][2] < min_intv) break;
			}
			ik.x[0] = ok[c][0]; ik.x[1] = ok[c][1]; ik.x[2] = ok[c][2]; ik.info = (uint64_t)(i + 1);
		} else {
			memv_push(curr, &ik);
			break;
		}
	}
	if (i == len) memv_push(curr, &ik);
	memv_reverse(curr); /* longest match first */
	ret = (int)curr->a[0].info;
	swap = curr; curr = prev; prev = swap;

	for (i = x - 1; i >= -1; --i) { /* backward sweep */
		c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
		for (j = 0, curr->n = 0; j < prev->n; ++j) {
			cso_mem_t *p = &prev->a[j];
			if (c >= 0) counted_extend_len(w, p, ok, 1, (int)(uint32_t)p->info - i);
			if (c < 0 || ok[c][2] < min_intv) {
				if (curr->n == 0) { /* no longer match survived this sweep */
					if (mem->n == 0 || (uint64_t)(i + 1) < mem->a[mem->n - 1].info >> 32) {
						t = *p; t.info |= (uint64_t)(i + 1) << 32;
						memv_push(mem, &t);
					}
				}
			} else if (curr->n == 0 || ok[c][2] != curr->a[curr->n - 1].x[2]) {
				t.x[0] = ok[c][0]; t.x[1] = ok[c][1]; t.x[2] = ok[c][2]; t.info = p->info;
				memv_push(curr, &t);
			}
		}
		if (curr->n == 0) break;
		swap = curr; curr = prev; prev = swap;
	}
	memv_reverse(mem);
	/* keep w->prev / w->curr pointing at their own storage regardless of the number of swaps */
	return ret;
}

/* ONE bwt_smem1a call for test harnesses (tests/emul/seed_emul.cpp resolves the calls the device kernels defer with it):
 * the SMEMs of the call, unfiltered, sorted by start, into out[cap]; returns their number (-1: cap too small), *ret = next pivot. */
int cso_smem1_call(const cso_index_t *idx, int len, const uint8_t *q, int x, uint64_t min_intv, cso_mem_t *out, int cap, int *ret)
{
	work_t w;
	int r, n;
	memset(&w, 0, sizeof w);
	w.idx = idx;
	r = smem1(&w, len, q, x, min_intv);
	if (ret) *ret = r;
	n = (int)w.mem1.n;
	if (n <= cap) memcpy(out, w.mem1.a, (size_t)n * sizeof(cso_mem_t)); else n = -1;
	free(w.prev.a); free(w.curr.a); free(w.mem1.a);
	return n;
}

/* bwt_seed_strategy1 (bwt.c:358-379) */
static int seed_strategy1(work_t *w, int len, const uint8_t *q, int x, int min_len, int max_intv, cso_mem_t *mem)
{
	cso_mem_t ik;
	uint64_t ok[4][3];
	int i, c;
	memset(mem, 0, sizeof(*mem));
	if (q[x] > 3) return x + 1;
	set_intv(w->idx, q[x], &ik);
	for (i = x + 1; i < len; ++i) {
		if (q[i] < 4) {
			c = 3 - q[i];
			counted_extend(w, &ik, ok, 0);
			if (ok[c][2] < (uint64_t)max_intv && i - x >= min_len) {
				mem->x[0] = ok[c][0]; mem->x[1] = ok[c][1]; mem->x[2] = ok[c][2];
				mem->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
				return i + 1;
			}
			ik.x[0] = ok[c][0]; ik.x[1] = ok[c][1]; ik.x[2] = ok[c][2];
		} else return i + 1;
	}
	return len;
}

static int mem_cmp(const void *a, const void *b)
{
	uint64_t x = ((const cso_mem_t*)a)->info, y = ((const cso_mem_t*)b)->info;
	return (x > y) - (x < y);
}

/* mem_collect_intv (bwamem.c:218-272) == seeding block of seed_and_extend (comp_seed.cpp:2255-2302) */
static void collect_intv(work_t *w, const cso_opt_t *opt, int len, const uint8_t *seq, memv_t *out)
{
	int x = 0;
	size_t i, k, old_n;
	out->n = 0;
	w->round = 0;
	while (x < len) {
		if (seq[x] < 4) {
			x = smem1(w, len, seq, x, 1);
			for (i = 0; i < w->mem1.n; ++i) {
				cso_mem_t *p = &w->mem1.a[i];
				if ((int)((uint32_t)p->info - (uint32_t)(p->info >> 32)) >= opt->min_seed_len) memv_push(out, p);
			}
		} else ++x;
	}
	w->round = 1;
	old_n = out->n;
	for (k = 0; k < old_n; ++k) {
		cso_mem_t p = out->a[k];
		int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
		if (end - start < opt->split_len || p.x[2] > (uint64_t)opt->split_width) continue;
		smem1(w, len, seq, (start + end) >> 1, p.x[2] + 1);
		for (i = 0; i < w->mem1.n; ++i) {
			cso_mem_t *m = &w->mem1.a[i];
			if ((int)((uint32_t)m->info - (uint32_t)(m->info >> 32)) >= opt->min_seed_len) memv_push(out, m);
		}
	}
	w->round = 2;
	if (opt->max_mem_intv > 0) {
		x = 0;
		while (x < len) {
			if (seq[x] < 4) {
				cso_mem_t m;
				x = seed_strategy1(w, len, seq, x, opt->min_seed_len, opt->max_mem_intv, &m);
				if (m.x[2] > 0) memv_push(out, &m);
			} else ++x;
		}
	}
	qsort(out->a, out->n, sizeof(cso_mem_t), mem_cmp);
}

typedef struct {
	int n_reads;
	uint32_t *mem_off, *seed_off;
	cso_mem_t *mems; uint64_t n_mems;
	int64_t *rbeg; uint64_t n_seeds;
	double seconds;
	int64_t cnt[CSO_N_CNT];
} result_t;

typedef struct { memv_t mems; int64_t *rbeg; size_t n_rbeg; } read_out_t;

typedef struct {
	const cso_index_t *idx; const cso_opt_t *opt;
	int n_reads; const uint8_t *bases; const uint32_t *off;
	int *next; read_out_t *out;
	int64_t cnt[CSO_N_CNT];
} job_t;

static void *seed_worker(void *arg)
{
	job_t *jb = (job_t*)arg;
	work_t w;
	memset(&w, 0, sizeof w);
	w.idx = jb->idx;
	for (;;) {
		int s = __atomic_fetch_add(jb->next, 256, __ATOMIC_RELAXED), e, r;
		if (s >= jb->n_reads) break;
		e = s + 256 < jb->n_reads ? s + 256 : jb->n_reads;
		for (r = s; r < e; ++r) {
			read_out_t *o = &jb->out[r];
			int len = (int)(jb->off[r + 1] - jb->off[r]);
			size_t i, ns = 0, cap;
			collect_intv(&w, jb->opt, len, jb->bases + jb->off[r], &o->mems);
			/* seed expansion, bwamem.c:386-399 == comp_seed.cpp:2309-2326: emission order is the contract */
			cap = 16;
			o->rbeg = (int64_t*)malloc(cap * 8);
			for (i = 0; i < o->mems.n; ++i) {
				const cso_mem_t *p = &o->mems.a[i];
				uint64_t step = p->x[2] > (uint64_t)jb->opt->max_occ ? p->x[2] / (uint64_t)jb->opt->max_occ : 1, k;
				int count;
				for (k = 0, count = 0; k < p->x[2] && count < jb->opt->max_occ; k += step, ++count) {
					if (ns == cap) { cap <<= 1; o->rbeg = (int64_t*)realloc(o->rbeg, cap * 8); }
					o->rbeg[ns++] = (int64_t)sa_lookup(jb->idx, p->x[0] + k, &w.cnt[CSO_CNT_LF]);
				}
			}
			o->n_rbeg = ns;
			w.cnt[CSO_CNT_SA] += (int64_t)ns;
			w.cnt[CSO_CNT_MEM] += (int64_t)o->mems.n;
		}
	}
	memcpy(jb->cnt, w.cnt, sizeof w.cnt);
	free(w.prev.a); free(w.curr.a); free(w.mem1.a);
	return 0;
}

void *cso_seed(const cso_index_t *idx, int n_threads, int n_reads, const uint8_t *bases, const uint32_t *off, const cso_opt_t *opt)
{
	int t, r, next = 0;
	struct timespec t0, t1;
	if (n_threads < 1) n_threads = 1;
	read_out_t *out = (read_out_t*)calloc((size_t)n_reads + 1, sizeof(read_out_t));
	job_t *jobs = (job_t*)calloc(n_threads, sizeof(job_t));
	pthread_t *th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (t = 0; t < n_threads; ++t) {
		jobs[t].idx = idx; jobs[t].opt = opt; jobs[t].n_reads = n_reads; jobs[t].bases = bases; jobs[t].off = off;
		jobs[t].next = &next; jobs[t].out = out;
		pthread_create(&th[t], 0, seed_worker, &jobs[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	result_t *res = (result_t*)calloc(1, sizeof(result_t));
	res->n_reads = n_reads;
	res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	res->mem_off = (uint32_t*)calloc((size_t)n_reads + 1, 4);
	res->seed_off = (uint32_t*)calloc((size_t)n_reads + 1, 4);
	for (r = 0; r < n_reads; ++r) {
		res->mem_off[r + 1] = res->mem_off[r] + (uint32_t)out[r].mems.n;
		res->seed_off[r + 1] = res->seed_off[r] + (uint32_t)out[r].n_rbeg;
	}
	res->n_mems = res->mem_off[n_reads]; res->n_seeds = res->seed_off[n_reads];
	res->mems = (cso_mem_t*)malloc((res->n_mems + 1) * sizeof(cso_mem_t));
	res->rbeg = (int64_t*)malloc((res->n_seeds + 1) * 8);
	for (r = 0; r < n_reads; ++r) {
		if (out[r].mems.n) memcpy(res->mems + res->mem_off[r], out[r].mems.a, out[r].mems.n * sizeof(cso_mem_t));
		if (out[r].n_rbeg) memcpy(res->rbeg + res->seed_off[r], out[r].rbeg, out[r].n_rbeg * 8);
		free(out[r].mems.a); free(out[r].rbeg);
	}
	for (t = 0; t < n_threads; ++t) { int i; for (i = 0; i < CSO_N_CNT; ++i) res->cnt[i] += jobs[t].cnt[i]; }
	free(out); free(jobs); free(th);
	return res;
}

uint64_t cso_result_n_mems(void *r) { return ((result_t*)r)->n_mems; }
uint64_t cso_result_n_seeds(void *r) { return ((result_t*)r)->n_seeds; }
double cso_result_seconds(void *r) { return ((result_t*)r)->seconds; }
void cso_result_counters(void *r, int64_t *out) { memcpy(out, ((result_t*)r)->cnt, sizeof(int64_t) * CSO_N_CNT); }
void cso_result_copy(void *rr, uint32_t *mem_off, uint64_t *mems, uint32_t *seed_off, int64_t *rbeg)
{
	result_t *r = (result_t*)rr;
	memcpy(mem_off, r->mem_off, ((size_t)r->n_reads + 1) * 4);
	memcpy(seed_off, r->seed_off, ((size_t)r->n_reads + 1) * 4);
	memcpy(mems, r->mems, r->n_mems * sizeof(cso_mem_t));
	memcpy(rbeg, r->rbeg, r->n_seeds * 8);
}
void cso_result_free(void *rr)
{
	result_t *r = (result_t*)rr;
	free(r->mem_off); free(r->seed_off); free(r->mems); free(r->rbeg); free(r);
}

/* ------------------------------------------------------------------------------------------------
 * Banded Smith-Waterman extension: ksw_extend2 (bwalib/ksw.c:380-479), which BandedPairWiseSW::scalarBandedSWA
 * (mapping/bandedSWA.cpp:118-237) repeats with m = 5.  Written from the recurrences the reference documents at ksw.c:423-428:
 *   H(i,j) = max{M(i,j), E(i,j), F(i,j)},  M(i,j) = H(i-1,j-1) ? H(i-1,j-1) + S(i,j) : 0
 *   E(i+1,j) = max{M(i,j) - o_del - e_del, E(i,j) - e_del, 0},  F(i,j+1) = max{M(i,j) - o_ins - e_ins, F(i,j) - e_ins, 0}
 * over a band that is re-cut to the non-zero cells after every row (ksw.c:463-468).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int32_t h, e; } cso_eh_t;

static int cso_bsw_one(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                       int w, int end_bonus, int zdrop, int h0, int *qle, int *tle, int *gtle, int *gscore_out, int *max_off_out, uint64_t *cells)
{
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	cso_eh_t *row = (cso_eh_t*)calloc((size_t)qlen + 1, sizeof(cso_eh_t));   /* row[j] = { H(i-1, j-1), E(i, j) } (ksw.c:393) */
	int i, j, k, best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0, lo = 0, hi = qlen, mmax = 0;
	uint64_t ncell = 0;
	/* H(-1, j): the query consumed as an insertion (ksw.c:395-398) */
	row[0].h = h0; row[1].h = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && row[j - 1].h > e_ins; ++j) row[j].h = row[j - 1].h - e_ins;
	/* the band cannot usefully be wider than the longest gap the best possible score pays for (ksw.c:400-408) */
	for (k = 0; k < 25; ++k) mmax = mmax > mat[k] ? mmax : mat[k];
	{
		int gi = (int)((double)(qlen * mmax + end_bonus - o_ins) / e_ins + 1.), gd = (int)((double)(qlen * mmax + end_bonus - o_del) / e_del + 1.);
		if (gi < 1) gi = 1;
		if (gd < 1) gd = 1;
		if (w > gi) w = gi;
		if (w > gd) w = gd;
	}
	for (i = 0; i < tlen; ++i) {
		const int8_t *s = mat + 5 * target[i];
		int f = 0, left, rowmax = 0, rowmax_j = -1;
		if (lo < i - w) lo = i - w;
		if (hi > i + w + 1) hi = i + w + 1;
		if (hi > qlen) hi = qlen;
		left = 0;                                             /* H(i, lo-1): the target consumed as a deletion, only in column -1 (ksw.c:416-419) */
		if (lo == 0) { left = h0 - (o_del + e_del * (i + 1)); if (left < 0) left = 0; }
		for (j = lo; j < hi; ++j) {
			const int diag = row[j].h, e = row[j].e;
			const int M = diag ? diag + s[query[j]] : 0;       /* ksw.c:432 */
			int h = M > e ? M : e, t;
			if (f > h) h = f;
			row[j].h = left; left = h;                        /* row[j].h becomes H(i, j-1) for the next row */
			if (h >= rowmax) { rowmax = h; rowmax_j = j; }    /* the LAST column that reaches the row maximum (ksw.c:437-438) */
			t = M - oe_del; if (t < 0) t = 0;
			row[j].e = e - e_del > t ? e - e_del : t;
			t = M - oe_ins; if (t < 0) t = 0;
			f = f - e_ins > t ? f - e_ins : t;
		}
		if (hi > lo) ncell += (uint64_t)(hi - lo);
		row[hi].h = left; row[hi].e = 0;
		if (j == qlen) {                                     /* the row reached the end of the query (ksw.c:449-452): >= keeps the LAST such row */
			if (left >= gscore) { best_ie = i; gscore = left; }
		}
		if (rowmax == 0) break;
		if (rowmax > best) {
			const int d = rowmax_j > i ? rowmax_j - i : i - rowmax_j;
			best = rowmax; best_i = i; best_j = rowmax_j;
			if (d > max_off) max_off = d;
		} else if (zdrop > 0) {
			const int di = i - best_i, dj = rowmax_j - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
			else if (best - rowmax - (dj - di) * e_ins > zdrop) break;
		}
		for (j = lo; j < hi && row[j].h == 0 && row[j].e == 0; ++j) {}
		lo = j;
		for (j = hi; j >= lo && row[j].h == 0 && row[j].e == 0; --j) {}
		hi = j + 2 < qlen ? j + 2 : qlen;
	}
	free(row);
	*cells += ncell;
	*qle = best_j + 1; *tle = best_i + 1; *gtle = best_ie + 1; *gscore_out = gscore; *max_off_out = max_off;
	return best;
}

typedef struct { int32_t *pairs; const uint8_t *ref, *qer; int n, w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus; const int8_t *mat; int t, nt; uint64_t cells; } cso_bsw_job_t;

static void *cso_bsw_worker(void *arg)
{
	cso_bsw_job_t *J = (cso_bsw_job_t*)arg;
	const int64_t a = (int64_t)J->n * J->t / J->nt, b = (int64_t)J->n * (J->t + 1) / J->nt;
	for (int64_t i = a; i < b; ++i) {
		int32_t *p = J->pairs + 14 * i;   /* SeqPair: idr idq id len1 len2 h0 seqid regid score tle gtle qle gscore max_off */
		int qle, tle, gtle, gscore, max_off;
		p[8] = cso_bsw_one(p[4], J->qer + p[1], p[3], J->ref + p[0], J->mat, J->o_del, J->e_del, J->o_ins, J->e_ins, J->w, J->end_bonus, J->zdrop, p[5],
		                   &qle, &tle, &gtle, &gscore, &max_off, &J->cells);
		p[9] = tle; p[10] = gtle; p[11] = qle; p[12] = gscore; p[13] = max_off;
	}
	return NULL;
}

uint64_t cso_bsw_extend(int32_t *pairs, const uint8_t *seq_buf_ref, const uint8_t *seq_buf_qer, int n_pairs, int w,
                        int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus, const int8_t *mat, int n_threads)
{
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 256) n_threads = 256;
	cso_bsw_job_t *J = (cso_bsw_job_t*)calloc(n_threads, sizeof(cso_bsw_job_t));
	pthread_t *th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
	uint64_t cells = 0;
	for (int t = 0; t < n_threads; ++t) {
		cso_bsw_job_t j = { pairs, seq_buf_ref, seq_buf_qer, n_pairs, w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, mat, t, n_threads, 0 };
		J[t] = j;
		pthread_create(&th[t], NULL, cso_bsw_worker, &J[t]);
	}
	for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); cells += J[t].cells; }
	free(J); free(th);
	return cells;
}
