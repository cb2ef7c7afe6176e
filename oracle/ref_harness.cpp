// TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product path.
//
// Thin driver around the UNMODIFIED reference objects (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libcsref.so).  It calls the reference's own seeding functions
// and returns their output in the flat result layout shared with the oracle and the CUDA path:
//
//   mode 0 ("bwamem"):   bwt_smem1 / bwt_seed_strategy1 / bwt_sa   (FM_index/bwt.h:104,123-126),
//                        driven exactly as mem_collect_intv does (mapping/bwamem.c:218-272) and
//                        expanded to seeds as mem_chain does (mapping/bwamem.c:386-399).
//   mode 1 ("CompSeed"): collect_mem_with_sst / tem_forward_sst   (mapping/comp_seed.cpp:67,141),
//                        driven exactly as seed_and_extend's seeding + SAL blocks do
//                        (mapping/comp_seed.cpp:2254-2346), one SST pair per thread, cleared per
//                        512-read block.
//
// Only the ~60 driver lines are restated here; every FM-index / SST operation is the reference's.

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <thread>
#include <atomic>
#include <chrono>

#include "FM_index/bwt.h"
#include "FM_index/bntseq.h"
#include "mapping/comp_seed.h"
#include "mapping/bandedSWA.h"
extern "C" {
#include "bwalib/ksw.h"
}

thread_aux_t tprof; // comp_seed.cpp:22 expects the driver to define it (main.cpp:15)

extern int collect_mem_with_sst(const uint8_t *seq, int len, int pivot, int min_hits, thread_aux_t &aux);
extern int tem_forward_sst(const mem_opt_t *opt, const uint8_t *seq, int len, int start, bwtintv_t *mem, thread_aux_t &aux);
// chaining and chain filtering (mapping/comp_seed.cpp:241-354); mem_chain_v is private to that file (comp_seed.cpp:176)
typedef struct { size_t n, m; mem_chain_t *a; } mem_chain_v;
extern mem_chain_v mem_chain(const mem_opt_t *opt, const bntseq_t *bns, int len, const std::vector<bwtintv_t> &mem, const std::vector<mem_seed_t> &seed);
extern int mem_chain_flt(const mem_opt_t *opt, int n_chn, mem_chain_t *a);

extern "C" {

typedef struct {
	int32_t min_seed_len;   // -k
	float   split_factor;   // -r
	int32_t split_width;    // -s
	int32_t max_mem_intv;   // -y
	int32_t max_occ;        // -c
} csref_opt_t;

typedef struct {
	bwt_t *bwt;
	int owns; // 1: loaded from files (bwt_destroy), 0: arrays borrowed from the caller
} csref_index_t;

typedef struct {
	int n_reads;
	std::vector<uint32_t> *mem_off, *seed_off;
	std::vector<bwtintv_t> *mems;
	std::vector<int64_t> *rbeg;
	double seconds;            // wall time of seeding + SA resolution
	int64_t counters[4];       // ext_queries, ext_calls, sal_queries, sal_calls (mode 1; mode 0: queries only where known)
} csref_result_t;

void *csref_index_load(const char *prefix)
{
	std::string p(prefix);
	csref_index_t *h = (csref_index_t*)calloc(1, sizeof(csref_index_t));
	h->bwt = bwt_restore_bwt((p + ".bwt").c_str());
	bwt_restore_sa((p + ".sa").c_str(), h->bwt);
	h->owns = 1;
	return h;
}

void *csref_index_from_arrays(uint64_t primary, const uint64_t *L2, uint64_t seq_len, uint32_t *bwt, uint64_t bwt_size,
                              uint64_t *sa, uint64_t n_sa, int sa_intv)
{
	csref_index_t *h = (csref_index_t*)calloc(1, sizeof(csref_index_t));
	bwt_t *b = (bwt_t*)calloc(1, sizeof(bwt_t));
	b->primary = primary;
	for (int i = 0; i < 5; ++i) b->L2[i] = L2[i];
	b->seq_len = seq_len; b->bwt_size = bwt_size; b->bwt = bwt;
	b->sa = sa; b->n_sa = n_sa; b->sa_intv = sa_intv;
	bwt_gen_cnt_table(b);
	h->bwt = b; h->owns = 0;
	return h;
}

// out[0..10] = primary, L2[0..4], seq_len, bwt_size, n_sa, sa_intv
void csref_index_info(void *hh, uint64_t *out)
{
	bwt_t *b = ((csref_index_t*)hh)->bwt;
	out[0] = b->primary;
	for (int i = 0; i < 5; ++i) out[1 + i] = b->L2[i];
	out[6] = b->seq_len; out[7] = b->bwt_size; out[8] = b->n_sa; out[9] = b->sa_intv;
}
const uint32_t *csref_index_bwt(void *hh) { return ((csref_index_t*)hh)->bwt->bwt; }
const uint64_t *csref_index_sa(void *hh) { return ((csref_index_t*)hh)->bwt->sa; }

void csref_index_free(void *hh)
{
	csref_index_t *h = (csref_index_t*)hh;
	if (h->owns) bwt_destroy(h->bwt); else free(h->bwt);
	free(h);
}

// Primitive probes (unit-level parity): the reference's own bwt_occ4 / bwt_extend / bwt_sa.
void csref_occ4(void *hh, int n, const uint64_t *k, uint64_t *cnt /*n*4*/)
{
	bwt_t *b = ((csref_index_t*)hh)->bwt;
	for (int i = 0; i < n; ++i) bwt_occ4(b, k[i], cnt + 4 * (size_t)i);
}
void csref_extend(void *hh, int n, const uint64_t *ik /*n*3*/, const int *is_back, uint64_t *ok /*n*4*3*/)
{
	bwt_t *b = ((csref_index_t*)hh)->bwt;
	for (int i = 0; i < n; ++i) {
		bwtintv_t in, out[4];
		in.x[0] = ik[3*i]; in.x[1] = ik[3*i+1]; in.x[2] = ik[3*i+2]; in.info = 0;
		bwt_extend(b, &in, out, is_back[i]);
		for (int c = 0; c < 4; ++c) for (int j = 0; j < 3; ++j) ok[(size_t)i*12 + c*3 + j] = out[c].x[j];
	}
}
void csref_sa(void *hh, int n, const uint64_t *k, uint64_t *out)
{
	bwt_t *b = ((csref_index_t*)hh)->bwt;
	for (int i = 0; i < n; ++i) out[i] = bwt_sa(b, k[i]);
}

struct read_out_t { std::vector<bwtintv_t> mems; std::vector<int64_t> rbeg; };

// mapping/bwamem.c:218-272 with the public bwt.c twins of smem1_profile / seed_strategy_profile
static void collect_bwamem(const bwt_t *bwt, const csref_opt_t *opt, int len, const uint8_t *seq,
                           bwtintv_v *mem1, bwtintv_v *tmpv[2], std::vector<bwtintv_t> &mem)
{
	int x = 0;
	int split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
	mem.clear();
	while (x < len) {
		if (seq[x] < 4) {
			x = bwt_smem1(bwt, len, seq, x, 1, mem1, tmpv);
			for (size_t i = 0; i < mem1->n; ++i) {
				bwtintv_t *p = &mem1->a[i];
				int slen = (uint32_t)p->info - (p->info >> 32);
				if (slen >= opt->min_seed_len) mem.push_back(*p);
			}
		} else ++x;
	}
	size_t old_n = mem.size();
	for (size_t k = 0; k < old_n; ++k) {
		bwtintv_t p = mem[k];
		int start = p.info >> 32, end = (int32_t)p.info;
		if (end - start < split_len || p.x[2] > (uint64_t)opt->split_width) continue;
		bwt_smem1(bwt, len, seq, (start + end) >> 1, p.x[2] + 1, mem1, tmpv);
		for (size_t i = 0; i < mem1->n; ++i)
			if ((uint32_t)mem1->a[i].info - (mem1->a[i].info >> 32) >= (uint32_t)opt->min_seed_len)
				mem.push_back(mem1->a[i]);
	}
	if (opt->max_mem_intv > 0) {
		x = 0;
		while (x < len) {
			if (seq[x] < 4) {
				bwtintv_t m;
				x = bwt_seed_strategy1(bwt, len, seq, x, opt->min_seed_len, opt->max_mem_intv, &m);
				if (m.x[2] > 0) mem.push_back(m);
			} else ++x;
		}
	}
	std::sort(mem.begin(), mem.end(), [](const bwtintv_t &a, const bwtintv_t &b) { return a.info < b.info; });
}

static void worker_bwamem(const bwt_t *bwt, const csref_opt_t *opt, const uint8_t *bases, const uint32_t *off,
                          int n_reads, std::atomic<int> *next, read_out_t *out)
{
	bwtintv_v mem1 = {0, 0, 0}, t0 = {0, 0, 0}, t1 = {0, 0, 0};
	bwtintv_v *tmpv[2] = { &t0, &t1 };
	const int chunk = 256;
	for (;;) {
		int s = next->fetch_add(chunk);
		if (s >= n_reads) break;
		int e = std::min(n_reads, s + chunk);
		for (int r = s; r < e; ++r) {
			int len = off[r + 1] - off[r];
			const uint8_t *seq = bases + off[r];
			collect_bwamem(bwt, opt, len, seq, &mem1, tmpv, out[r].mems);
			out[r].rbeg.clear();
			for (const bwtintv_t &p : out[r].mems) { // mapping/bwamem.c:386-399
				int64_t k; int count;
				int step = p.x[2] > (uint64_t)opt->max_occ ? p.x[2] / opt->max_occ : 1;
				for (k = count = 0; k < (int64_t)p.x[2] && count < opt->max_occ; k += step, ++count)
					out[r].rbeg.push_back((int64_t)bwt_sa(bwt, p.x[0] + k));
			}
		}
	}
	free(mem1.a); free(t0.a); free(t1.a);
}

static inline int mem_beg(const bwtintv_t &a) { return a.info >> 32; }
static inline int mem_end(const bwtintv_t &a) { return (int)a.info; }
static inline int mem_len(const bwtintv_t &a) { return mem_end(a) - mem_beg(a); }

// mapping/comp_seed.cpp:2254-2346, one 512-read block at a time
static void worker_compseed(const bwt_t *bwt, const csref_opt_t *o, const uint8_t *bases, const uint32_t *off,
                            int n_reads, std::atomic<int> *next, read_out_t *out, int64_t *counters)
{
	thread_aux_t *auxp = new thread_aux_t();
	thread_aux_t &aux = *auxp;
	aux.forward_sst = new SST(bwt);
	aux.backward_sst = new SST(bwt);
	mem_opt_t *opt = mem_opt_init();
	opt->min_seed_len = o->min_seed_len; opt->split_factor = o->split_factor; opt->split_width = o->split_width;
	opt->max_mem_intv = o->max_mem_intv; opt->max_occ = o->max_occ;
	for (;;) {
		int s = next->fetch_add(BATCH_SIZE);
		if (s >= n_reads) break;
		int n = std::min(n_reads, s + BATCH_SIZE) - s;
		aux.forward_sst->clear(); aux.backward_sst->clear();
		for (int r = 0; r < n; r++) {
			int l_seq = off[s + r + 1] - off[s + r];
			const uint8_t *seq = bases + off[s + r];
			std::vector<bwtintv_t> &match = aux.match[r]; match.clear();
			for (int j = 0; j < l_seq; ) {
				j = collect_mem_with_sst(seq, l_seq, j, 1, aux);
				for (const auto &m : aux.super_mem)
					if (mem_len(m) >= opt->min_seed_len) match.push_back(m);
			}
			int old_n = (int)match.size();
			for (int j = 0; j < old_n; j++) {
				const auto p = match[j];
				int beg = mem_beg(p), end = mem_end(p);
				if (end - beg < (int)(1.0 * opt->min_seed_len * opt->split_factor + .499) or p.x[2] > (uint64_t)opt->split_width) continue;
				collect_mem_with_sst(seq, l_seq, (beg + end) / 2, p.x[2] + 1, aux);
				for (const auto &m : aux.super_mem)
					if (mem_len(m) >= opt->min_seed_len) match.push_back(m);
			}
			if (opt->max_mem_intv > 0) {
				for (int j = 0; j < l_seq; ) {
					if (seq[j] < 4) {
						bwtintv_t m;
						j = tem_forward_sst(opt, seq, l_seq, j, &m, aux);
						if (m.x[2] > 0) match.push_back(m);
					} else j++;
				}
			}
			std::sort(match.begin(), match.end(), [](const bwtintv_t &a, const bwtintv_t &b) { return a.info < b.info; });
		}
		aux.bwt_call_times += aux.forward_sst->bwt_call + aux.backward_sst->bwt_call;

		auto &unique_sal = aux.unique_sal; unique_sal.clear();
		for (int r = 0; r < n; r++) {
			const auto &mem = aux.match[r];
			auto &seed = aux.seed[r]; seed.clear();
			for (const auto &m : mem) {
				uint64_t step = m.x[2] > (uint64_t)opt->max_occ ? m.x[2] / opt->max_occ : 1;
				for (uint64_t k = 0, count = 0; k < m.x[2] && count < (uint64_t)opt->max_occ; k += step, count++) {
					mem_seed_t sd; memset(&sd, 0, sizeof(sd));
					sd.qbeg = mem_beg(m);
					sd.score = sd.len = mem_len(m);
					sd.rbeg = m.x[0] + k;
					seed.push_back(sd);
					unique_sal.emplace_back(sal_request_t(m.x[0] + k));
					aux.sal_query_times++;
				}
			}
		}
		std::sort(unique_sal.begin(), unique_sal.end());
		int _size = 0;
		for (size_t i = 0; i < unique_sal.size(); i++)
			if (i == 0 or unique_sal[i-1].que_location != unique_sal[i].que_location) unique_sal[_size++] = unique_sal[i];
		unique_sal.resize(_size);
		for (int r = 0; r < n; r++) {
			for (auto &sd : aux.seed[r]) {
				auto k = std::lower_bound(unique_sal.begin(), unique_sal.end(), sal_request_t(sd.rbeg));
				if (k->coordinate == (uint64_t)-1) {
					k->coordinate = bwt_sa(bwt, sd.rbeg);
					aux.sal_call_times++;
				}
				sd.rbeg = k->coordinate;
			}
		}
		for (int r = 0; r < n; r++) {
			out[s + r].mems = aux.match[r];
			out[s + r].rbeg.clear();
			for (auto &sd : aux.seed[r]) out[s + r].rbeg.push_back(sd.rbeg);
		}
	}
	counters[0] = aux.bwt_query_times; counters[1] = aux.bwt_call_times;
	counters[2] = aux.sal_query_times; counters[3] = aux.sal_call_times;
	delete aux.forward_sst; delete aux.backward_sst;
	delete auxp;
	free(opt);
}

// ---------------------------------------------------------------------------------------------
// Chains: the reference's own mem_chain + mem_chain_flt (comp_seed.cpp:241-354) on given mems and seeds.
// Contigs: n_seqs lengths (a bntseq_t with just the fields bns_intv2rid reads, bntseq.c:354-378).  Output, flat:
//   chain_off[n_reads+1]; per chain pos, rid, w, kept | is_alt << 8, n seeds; seeds of all chains in chain order
//   (rbeg, qbeg, len); per read frac_rep (float) of its chains (0 if it has none).
// ---------------------------------------------------------------------------------------------
typedef struct {
	int32_t w, max_chain_gap, min_chain_weight, max_chain_extend;
	float mask_level, drop_ratio;
} csref_chain_opt_t;

typedef struct {
	std::vector<uint32_t> chain_off;
	std::vector<int64_t> pos; std::vector<int32_t> rid, w, kept, n;
	std::vector<int64_t> s_rbeg; std::vector<int32_t> s_qbeg, s_len;
	std::vector<float> frac_rep;
} csref_chains_t;

void *csref_chain(int n_reads, const uint32_t *read_off, const uint32_t *mem_off, const uint64_t *mems /*n*4*/, const uint32_t *seed_off,
                  const int64_t *rbeg, const csref_opt_t *so, const csref_chain_opt_t *co, int n_seqs, const int32_t *seq_len, const uint8_t *is_alt)
{
	bntseq_t bns; memset(&bns, 0, sizeof bns);
	std::vector<bntann1_t> anns(n_seqs);
	int64_t o = 0;
	for (int i = 0; i < n_seqs; ++i) { memset(&anns[i], 0, sizeof(bntann1_t)); anns[i].offset = o; anns[i].len = seq_len[i]; anns[i].is_alt = is_alt ? is_alt[i] : 0; o += seq_len[i]; }
	bns.l_pac = o; bns.n_seqs = n_seqs; bns.anns = anns.data();
	mem_opt_t *opt = mem_opt_init();
	opt->min_seed_len = so->min_seed_len; opt->split_factor = so->split_factor; opt->split_width = so->split_width;
	opt->max_mem_intv = so->max_mem_intv; opt->max_occ = so->max_occ;
	opt->w = co->w; opt->max_chain_gap = co->max_chain_gap; opt->min_chain_weight = co->min_chain_weight; opt->max_chain_extend = co->max_chain_extend;
	opt->mask_level = co->mask_level; opt->drop_ratio = co->drop_ratio;
	csref_chains_t *out = new csref_chains_t();
	out->chain_off.assign(n_reads + 1, 0);
	out->frac_rep.assign(n_reads, 0.f);
	for (int r = 0; r < n_reads; ++r) {
		std::vector<bwtintv_t> mem(mem_off[r + 1] - mem_off[r]);
		for (size_t i = 0; i < mem.size(); ++i) { const uint64_t *p = mems + 4 * (size_t)(mem_off[r] + i); mem[i].x[0] = p[0]; mem[i].x[1] = p[1]; mem[i].x[2] = p[2]; mem[i].info = p[3]; }
		std::vector<mem_seed_t> seed; // expanded exactly as comp_seed.cpp:2309-2326 does (then resolved: rbeg given)
		size_t si = seed_off[r];
		for (const auto &m : mem) {
			uint64_t step = m.x[2] > (uint64_t)opt->max_occ ? m.x[2] / opt->max_occ : 1;
			for (uint64_t k = 0, count = 0; k < m.x[2] && count < (uint64_t)opt->max_occ; k += step, count++) {
				mem_seed_t sd; memset(&sd, 0, sizeof(sd));
				sd.qbeg = m.info >> 32; sd.score = sd.len = (int)m.info - (int)(m.info >> 32); sd.rbeg = rbeg[si++];
				seed.push_back(sd);
			}
		}
		mem_chain_v chn = mem_chain(opt, &bns, (int)(read_off[r + 1] - read_off[r]), mem, seed);
		chn.n = mem_chain_flt(opt, chn.n, chn.a);
		for (size_t c = 0; c < chn.n; ++c) {
			const mem_chain_t &ch = chn.a[c];
			out->pos.push_back(ch.pos); out->rid.push_back(ch.rid); out->w.push_back((int32_t)ch.w);
			out->kept.push_back((int32_t)ch.kept | ((int32_t)ch.is_alt << 8)); out->n.push_back(ch.n);
			for (int j = 0; j < ch.n; ++j) { out->s_rbeg.push_back(ch.seeds[j].rbeg); out->s_qbeg.push_back(ch.seeds[j].qbeg); out->s_len.push_back(ch.seeds[j].len); }
			out->frac_rep[r] = ch.frac_rep;
			free(ch.seeds);
		}
		free(chn.a);
		out->chain_off[r + 1] = (uint32_t)out->pos.size();
	}
	free(opt);
	return out;
}
uint64_t csref_chains_n(void *h) { return ((csref_chains_t*)h)->pos.size(); }
uint64_t csref_chains_n_seeds(void *h) { return ((csref_chains_t*)h)->s_rbeg.size(); }
void csref_chains_copy(void *h, uint32_t *chain_off, int64_t *pos, int32_t *rid, int32_t *w, int32_t *kept, int32_t *n,
                       int64_t *s_rbeg, int32_t *s_qbeg, int32_t *s_len, float *frac_rep)
{
	csref_chains_t *c = (csref_chains_t*)h;
	memcpy(chain_off, c->chain_off.data(), c->chain_off.size() * 4);
	if (!c->pos.empty()) {
		memcpy(pos, c->pos.data(), c->pos.size() * 8); memcpy(rid, c->rid.data(), c->rid.size() * 4); memcpy(w, c->w.data(), c->w.size() * 4);
		memcpy(kept, c->kept.data(), c->kept.size() * 4); memcpy(n, c->n.data(), c->n.size() * 4);
	}
	if (!c->s_rbeg.empty()) { memcpy(s_rbeg, c->s_rbeg.data(), c->s_rbeg.size() * 8); memcpy(s_qbeg, c->s_qbeg.data(), c->s_qbeg.size() * 4); memcpy(s_len, c->s_len.data(), c->s_len.size() * 4); }
	memcpy(frac_rep, c->frac_rep.data(), c->frac_rep.size() * 4);
}
void csref_chains_free(void *h) { delete (csref_chains_t*)h; }

// bases: nt4 codes (0..3, >3 ambiguous) concatenated; off: n_reads+1 offsets.
void *csref_seed(void *hh, int mode, int n_threads, int n_reads, const uint8_t *bases, const uint32_t *off, const csref_opt_t *opt)
{
	const bwt_t *bwt = ((csref_index_t*)hh)->bwt;
	if (n_threads < 1) n_threads = 1;
	std::vector<read_out_t> out(n_reads);
	std::atomic<int> next(0);
	std::vector<int64_t> cnt((size_t)n_threads * 4, 0);
	auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; ++t) {
		if (mode == 0) th.emplace_back(worker_bwamem, bwt, opt, bases, off, n_reads, &next, out.data());
		else th.emplace_back(worker_compseed, bwt, opt, bases, off, n_reads, &next, out.data(), cnt.data() + 4 * t);
	}
	for (auto &t : th) t.join();
	auto t1 = std::chrono::steady_clock::now();

	csref_result_t *res = new csref_result_t();
	res->n_reads = n_reads;
	res->seconds = std::chrono::duration<double>(t1 - t0).count();
	res->mem_off = new std::vector<uint32_t>(n_reads + 1, 0);
	res->seed_off = new std::vector<uint32_t>(n_reads + 1, 0);
	res->mems = new std::vector<bwtintv_t>();
	res->rbeg = new std::vector<int64_t>();
	for (int r = 0; r < n_reads; ++r) {
		res->mems->insert(res->mems->end(), out[r].mems.begin(), out[r].mems.end());
		res->rbeg->insert(res->rbeg->end(), out[r].rbeg.begin(), out[r].rbeg.end());
		(*res->mem_off)[r + 1] = res->mems->size();
		(*res->seed_off)[r + 1] = res->rbeg->size();
	}
	for (int i = 0; i < 4; ++i) { res->counters[i] = 0; for (int t = 0; t < n_threads; ++t) res->counters[i] += cnt[4*t+i]; }
	return res;
}

uint64_t csref_result_n_mems(void *r) { return ((csref_result_t*)r)->mems->size(); }
uint64_t csref_result_n_seeds(void *r) { return ((csref_result_t*)r)->rbeg->size(); }
double csref_result_seconds(void *r) { return ((csref_result_t*)r)->seconds; }
void csref_result_counters(void *r, int64_t *out) { memcpy(out, ((csref_result_t*)r)->counters, 4 * sizeof(int64_t)); }
void csref_result_copy(void *rr, uint32_t *mem_off, uint64_t *mems /*n_mems*4*/, uint32_t *seed_off, int64_t *rbeg)
{
	csref_result_t *r = (csref_result_t*)rr;
	memcpy(mem_off, r->mem_off->data(), r->mem_off->size() * 4);
	memcpy(seed_off, r->seed_off->data(), r->seed_off->size() * 4);
	if (!r->mems->empty()) memcpy(mems, r->mems->data(), r->mems->size() * sizeof(bwtintv_t));
	if (!r->rbeg->empty()) memcpy(rbeg, r->rbeg->data(), r->rbeg->size() * 8);
}
void csref_result_free(void *rr)
{
	csref_result_t *r = (csref_result_t*)rr;
	delete r->mem_off; delete r->seed_off; delete r->mems; delete r->rbeg; delete r;
}


// ---- banded Smith-Waterman extension: the reference's own batch entry points (mapping/bandedSWA.cpp) ----
// mode 0: scalarBandedSWAWrapper (:242-260, == ksw_extend2 per pair); mode 1: getScores8; mode 2: getScores16 (the SIMD twins
// mem_chain2aln_across_reads_V2 calls, comp_seed.cpp:1790,1859; they write into pairs[n .. roundup(n, SIMD width)), so the
// batch is copied into a padded array first, as the caller's arrays are over-allocated by MAX_LINE_LEN, comp_seed.cpp:1491).
// pairs: 14 int32 each == SeqPair (bandedSWA.h:91-99).  Returns seconds spent inside the reference call.
double csref_bsw(int32_t *pairs, const uint8_t *seq_buf_ref, const uint8_t *seq_buf_qer, int n_pairs, int w,
                 int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus, const int8_t *mat, int w_match, int w_mismatch, int mode)
{
	static_assert(sizeof(SeqPair) == 14 * sizeof(int32_t), "SeqPair layout");
	BandedPairWiseSW bsw(o_del, e_del, o_ins, e_ins, zdrop, end_bonus, mat, (int8_t)w_match, (int8_t)w_mismatch, 1);
	std::vector<SeqPair> buf((size_t)n_pairs + 2 * MAX_LINE_LEN);
	memcpy(buf.data(), pairs, (size_t)n_pairs * sizeof(SeqPair));
	const auto t0 = std::chrono::steady_clock::now();
	if (mode == 0) bsw.scalarBandedSWAWrapper(buf.data(), (uint8_t*)seq_buf_ref, (uint8_t*)seq_buf_qer, n_pairs, 1, w);
	else if (mode == 1) bsw.getScores8(buf.data(), (uint8_t*)seq_buf_ref, (uint8_t*)seq_buf_qer, n_pairs, 1, w);
	else bsw.getScores16(buf.data(), (uint8_t*)seq_buf_ref, (uint8_t*)seq_buf_qer, n_pairs, 1, w);
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	memcpy(pairs, buf.data(), (size_t)n_pairs * sizeof(SeqPair));
	return sec;
}

// ksw_extend2 itself (bwalib/ksw.c:380), the call of bwamem's mem_chain2aln (mapping/bwamem.c:720,746), for one pair
int csref_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                      int w, int end_bonus, int zdrop, int h0, int *out5 /* qle tle gtle gscore max_off */)
{
	return ksw_extend2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, out5, out5 + 1, out5 + 2, out5 + 3, out5 + 4);
}

} // extern "C"
