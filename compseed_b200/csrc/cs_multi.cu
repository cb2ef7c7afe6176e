// Multi-device batch pipeline behind the C-ABI (north_star: "kthread workers swapped for a
// pinned-buffer, multi-stream batch pipeline ... results gathered in input order on the host with no NCCL").
//
// What it replaces in the reference: kt_for(opt->n_threads, worker1, ...) over the reads of a -K batch
// (mapping/bwamem.c:1343, comp_seed.cpp:2541-2548) for the seeding part, and -- through the two read sets that may be
// in flight -- the overlap kt_pipeline gives between reading batch i+1 and processing batch i (fastmap.c:76-140,
// kthread.c:95-107).
//
// One index replica and one cs_ctx per device.  A read set is split into one CONTIGUOUS block of reads per device, in
// input order, each a multiple of 512 reads (the reference's reuse block BATCH_SIZE, comp_seed.h:36; neighbouring
// reordered reads stay on one GPU).  One host thread per device pipelines its block in batches through the slots of its
// ctx: submit (DMA straight out of the caller's page-locked reads, or slices of the caller's packed set), wait for the
// kernels, enqueue the result copy -- compact wire format, straight into the block's page-locked result arrays at the
// position the previous batches left -- and go on.  No host thread ever touches a result byte; nothing is exchanged
// between devices; "gathering in input order" is the accessor cs_multi_read (block by read index, batch by position).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <deque>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <algorithm>
#include "cs_internal.h"

#define CS_TRACE_COLS 8

namespace {

struct BlockBuf { // page-locked result arrays of one (set, device)
	cs_block_t pub;
	uint64_t *mem_base = nullptr, *seed_base = nullptr; uint32_t cap_batches = 0;
	uint32_t *mem_off = nullptr, *seed_off = nullptr; uint64_t cap_off = 0;
	cs_cmem_t *cmems = nullptr; uint64_t cap_mems = 0;
	uint32_t *rlo = nullptr; uint8_t *rhi = nullptr; uint64_t cap_seeds = 0;
	cs_chain_t *chains = nullptr; uint64_t cap_chains = 0;              // chaining mode: chains instead of cmems ...
	uint16_t *sq = nullptr, *sl = nullptr; uint64_t cap_sq = 0;          // ... and qbeg / len next to the positions
};

struct Job {
	int set;
	uint64_t n_reads;
	const uint8_t *bases; const uint64_t *packed; const uint32_t *nmask; const uint64_t *off;
	cs_seed_opt_t opt;
};

struct Dev {
	cs_multi *m; int k;                // position in the device list
	const cs_index *idx; cs_ctx *ctx;
	uint64_t cap_mems, cap_seeds;      // result capacities of a slot of ctx
	std::thread th;
	std::mutex mu; std::condition_variable cv;
	std::deque<Job> q; bool stop = false;
	BlockBuf buf[2];
	int rc[2]; char err[2][512];
	uint64_t n_mems[2], n_seeds[2]; cs_counters_t cnt[2];
	double host_s[2][3];               // seconds spent submitting / polling + enqueueing copies / idle
	double gpu_ms[2][4];               // summed over the batches of a set: see cs_multi_result_t.gpu_ms
	bool done[2];
	std::vector<float> trace[2];       // diagnostics: CS_TRACE_COLS floats per batch of the last run of a set (cs_multi_trace)
};

} // namespace

struct cs_multi {
	int n_dev;
	uint32_t batch_reads, max_read_len; int n_slots;
	uint32_t mems_per_read, seeds_per_read;
	cs_ctx_config_t cfg;
	bool chaining; cs_chain_opt_t copt;                     // cs_multi_set_chaining: contig table kept for ctx re-creation
	int64_t l_pac; std::vector<int64_t> c_off; std::vector<uint8_t> c_alt;
	std::vector<Dev*> dev;
	cs_block_t blocks[2][CS_MULTI_MAX_DEV];
	bool busy[2];
	std::chrono::steady_clock::time_point t0[2];
	uint64_t n_reads[2];
};

namespace {

int pinned_grow(void **p, uint64_t *cap, uint64_t need, size_t elt, uint64_t keep)
{ // page-locked buffer of at least `need` elements; the first `keep` survive
	if (need <= *cap) return CS_OK;
	const uint64_t ncap = std::max<uint64_t>(need + need / 4 + 1024, *cap * 2);
	void *np = nullptr;
	if (cudaMallocHost(&np, ncap * elt) != cudaSuccess) { cudaGetLastError(); return cs_set_err(CS_E_CUDA, "cudaMallocHost(%llu bytes) failed", (unsigned long long)(ncap * elt)); }
	if (*p) { if (keep) memcpy(np, *p, keep * elt); cudaFreeHost(*p); }
	*p = np; *cap = ncap;
	return CS_OK;
}

void block_bounds(uint64_t n_reads, int n_dev, int k, uint64_t *r0, uint64_t *r1)
{ // contiguous, in input order, whole multiples of the 512-read reuse block (comp_seed.h:36) except the last
	const uint64_t nb = (n_reads + 511) / 512;
	*r0 = std::min<uint64_t>(n_reads, (nb * (uint64_t)k / n_dev) * 512);
	*r1 = std::min<uint64_t>(n_reads, (nb * (uint64_t)(k + 1) / n_dev) * 512);
}

int make_ctx(Dev *d)
{
	cs_multi *m = d->m;
	cs_ctx_config_t cfg = m->cfg;
	cfg.compact_results = 1;
	if (d->ctx) { cs_ctx_free(d->ctx); d->ctx = nullptr; }
	d->ctx = cs_ctx_create_ex(d->idx, m->batch_reads, (uint64_t)m->batch_reads * m->max_read_len, m->max_read_len, d->cap_mems, d->cap_seeds, m->n_slots, &cfg);
	if (!d->ctx) return CS_E_CUDA;
	if (m->chaining) {
		cs_bns_view_t v; v.l_pac = m->l_pac; v.n_seqs = (int32_t)m->c_off.size(); v.offset = m->c_off.data(); v.is_alt = m->c_alt.empty() ? nullptr : m->c_alt.data();
		return cs_ctx_set_chaining(d->ctx, &v, &m->copt);
	}
	return CS_OK;
}

// One read set on one device: reads [r0, r1) in batches through the slots.  A Run is a set being pipelined; the worker keeps up to
// two of them going at once, so that the first batches of set i+1 are copied in and seeded while the last batches of set i are on
// their way out (otherwise the pipeline drains between sets: ~13 ms of 83 per 10 M-read set on the B200 box).
struct Run {
	Job j;
	BlockBuf *b;
	uint64_t r0 = 0, r1 = 0, n = 0; uint32_t nb = 0;
	uint32_t next = 0, done = 0, copied = 0;   // batches submitted / kernels finished and result copy enqueued / results on the host
	int rc = CS_OK;
};
struct Flight { Run *run; uint32_t bi; int slot; };

int run_prepare(Dev *d, Run *r)
{
	cs_multi *m = d->m;
	const Job &j = r->j;
	BlockBuf &b = d->buf[j.set];
	r->b = &b;
	block_bounds(j.n_reads, m->n_dev, d->k, &r->r0, &r->r1);
	const uint64_t n = r->r1 - r->r0, B = m->batch_reads;
	const uint32_t nb = (uint32_t)((n + B - 1) / B);
	r->n = n; r->nb = nb;
	int rc;
	memset(&b.pub, 0, sizeof b.pub);
	b.pub.r0 = r->r0; b.pub.r1 = r->r1; b.pub.batch_reads = m->batch_reads; b.pub.n_batches = nb; b.pub.device = d->idx->device;
	d->n_mems[j.set] = d->n_seeds[j.set] = 0; memset(&d->cnt[j.set], 0, sizeof(cs_counters_t));
	d->host_s[j.set][0] = d->host_s[j.set][1] = d->host_s[j.set][2] = 0;
	for (int q = 0; q < 4; ++q) d->gpu_ms[j.set][q] = 0;
	d->trace[j.set].assign((size_t)nb * CS_TRACE_COLS, 0.f);
	if (n == 0) return CS_OK;
	if (cs_use_device(d->idx->device) != CS_OK) return CS_E_CUDA;
	{
		uint64_t cb = b.cap_batches, co = b.cap_off;
		void *p;
		p = b.mem_base; if ((rc = pinned_grow(&p, &cb, nb + 1, 8, 0)) != CS_OK) return rc; b.mem_base = (uint64_t*)p;
		cb = b.cap_batches; p = b.seed_base; if ((rc = pinned_grow(&p, &cb, nb + 1, 8, 0)) != CS_OK) return rc; b.seed_base = (uint64_t*)p;
		b.cap_batches = (uint32_t)cb;
		p = b.mem_off; if ((rc = pinned_grow(&p, &co, (uint64_t)nb * (B + 1), 4, 0)) != CS_OK) return rc; b.mem_off = (uint32_t*)p;
		co = b.cap_off; p = b.seed_off; if ((rc = pinned_grow(&p, &co, (uint64_t)nb * (B + 1), 4, 0)) != CS_OK) return rc; b.seed_off = (uint32_t*)p;
		b.cap_off = co;
		uint64_t cm = b.cap_mems, cs = b.cap_seeds;
		if (m->chaining) {
			uint64_t cc = b.cap_chains, cq = b.cap_sq;
			p = b.chains; if ((rc = pinned_grow(&p, &cc, n * 4, sizeof(cs_chain_t), 0)) != CS_OK) return rc; b.chains = (cs_chain_t*)p; b.cap_chains = cc;
			p = b.sq; if ((rc = pinned_grow(&p, &cq, n * m->seeds_per_read, 2, 0)) != CS_OK) return rc; b.sq = (uint16_t*)p;
			cq = b.cap_sq; p = b.sl; if ((rc = pinned_grow(&p, &cq, n * m->seeds_per_read, 2, 0)) != CS_OK) return rc; b.sl = (uint16_t*)p; b.cap_sq = cq;
		} else {
			p = b.cmems; if ((rc = pinned_grow(&p, &cm, n * m->mems_per_read, sizeof(cs_cmem_t), 0)) != CS_OK) return rc; b.cmems = (cs_cmem_t*)p; b.cap_mems = cm;
		}
		p = b.rlo; if ((rc = pinned_grow(&p, &cs, n * m->seeds_per_read, 4, 0)) != CS_OK) return rc; b.rlo = (uint32_t*)p;
		cs = b.cap_seeds; p = b.rhi; if ((rc = pinned_grow(&p, &cs, n * m->seeds_per_read, 1, 0)) != CS_OK) return rc; b.rhi = (uint8_t*)p;
		b.cap_seeds = cs;
	}
	b.mem_base[0] = b.seed_base[0] = 0;
	return CS_OK;
}

void run_publish(Dev *d, Run *r)
{
	cs_multi *m = d->m;
	BlockBuf &b = *r->b;
	const int set = r->j.set;
	if (r->rc == CS_OK && r->n) {
		d->n_mems[set] = b.mem_base[r->nb]; d->n_seeds[set] = b.seed_base[r->nb];
		b.pub.mem_base = b.mem_base; b.pub.seed_base = b.seed_base; b.pub.mem_off = b.mem_off; b.pub.seed_off = b.seed_off;
		b.pub.cmems = m->chaining ? nullptr : b.cmems; b.pub.rbeg_lo = b.rlo; b.pub.rbeg_hi = b.rhi;
		b.pub.chains = m->chaining ? b.chains : nullptr; b.pub.qbeg = m->chaining ? b.sq : nullptr; b.pub.len = m->chaining ? b.sl : nullptr;
	}
	{
		std::lock_guard<std::mutex> lk(d->mu);
		d->rc[set] = r->rc;
		if (r->rc != CS_OK) { strncpy(d->err[set], cs_last_error(), sizeof d->err[set] - 1); d->err[set][sizeof d->err[set] - 1] = 0; }
		d->done[set] = true;
	}
	d->cv.notify_all();
}

// Event-driven: the thread never blocks on one thing while another is ready.  Whenever the kernels of the oldest unfinished batch
// are done, its result copy is enqueued at once (the copy engine must not wait for the host); whenever the oldest copy has landed,
// its slot is free for the next batch -- of the same set or of the next one; otherwise the thread naps for 20 us.
void worker(Dev *d)
{
	cs_multi *m = d->m;
	std::deque<Run*> runs;          // sets being pipelined, oldest first
	std::deque<Flight> fl;          // batches in flight in submission order; fl[0 .. enq) have their result copy enqueued
	size_t enq = 0;
	uint64_t seq = 0;               // batches submitted so far: batch number seq uses slot seq % n_slots
	const uint64_t B = m->batch_reads;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
	long long base_ns = 0;
	auto host_ms = [&]() -> float { return (float)((std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() - base_ns) * 1e-6); };
	auto trace_row = [&](Run *r, uint32_t bi) -> float* { return d->trace[r->j.set].data() + (size_t)bi * CS_TRACE_COLS; };
	auto retire = [&](Run *r) { // r is runs.front() and complete (or failed)
		run_publish(d, r);
		runs.pop_front();
		delete r;
	};
	// a failure: everything in flight is waited for and every active set fails with that code
	auto fail_all = [&](int rc) {
		for (size_t i = 0; i < fl.size(); ++i) {
			uint64_t a, b2;
			if (i < enq) cs_i_fetch_wait(d->ctx, fl[i].slot, nullptr, nullptr); else cs_i_finish(d->ctx, fl[i].slot, &a, &b2);
		}
		fl.clear(); enq = 0;
		while (!runs.empty()) { runs.front()->rc = rc; retire(runs.front()); }
	};
	// results of fl[0] have landed
	auto complete_front = [&](bool blocking) -> int {
		Flight f = fl.front();
		Run *r = f.run;
		const int set = r->j.set;
		if (!blocking) {
			const int pr = cs_i_poll(d->ctx, f.slot);
			if (pr <= 0) return pr;
		}
		float ms4[4] = {0, 0, 0, 0};
		const int rc = cs_i_fetch_wait(d->ctx, f.slot, &d->cnt[set], ms4);
		if (rc != CS_OK) return rc;
		for (int q = 0; q < 4; ++q) d->gpu_ms[set][q] += ms4[q];
		cs_i_slot_times(d->ctx, f.slot, trace_row(r, f.bi), nullptr);
		trace_row(r, f.bi)[7] = host_ms();
		fl.pop_front(); --enq;
		++r->copied;
		return 1;
	};
	for (;;) {
		{ // new sets: wait for one when idle, otherwise just look
			std::unique_lock<std::mutex> lk(d->mu);
			if (runs.empty()) d->cv.wait(lk, [&] { return d->stop || !d->q.empty(); });
			if (runs.empty() && d->q.empty()) return;   // stop
			while (!d->q.empty() && runs.size() < 2) {
				Run *r = new Run();
				r->j = d->q.front(); d->q.pop_front();
				runs.push_back(r);
				lk.unlock();
				r->rc = run_prepare(d, r);
				if (base_ns == 0 && d->ctx) { float t5[5]; cs_i_slot_times(d->ctx, 0, t5, &base_ns); }
				lk.lock();
			}
		}
		while (!runs.empty() && (runs.front()->rc != CS_OK || runs.front()->copied == runs.front()->nb)) {
			if (runs.front()->rc != CS_OK && !fl.empty()) { fail_all(runs.front()->rc); break; }
			retire(runs.front());
		}
		if (runs.empty()) continue;
		bool progress = false;
		int rc = CS_OK;
		// ---- submit: a slot is free once the results of its previous batch have landed ----
		while (fl.size() < (size_t)m->n_slots) {
			Run *r = nullptr;
			for (Run *c : runs) if (c->rc == CS_OK && c->next < c->nb) { r = c; break; }
			if (!r) break;
			const auto t = now();
			const uint32_t bi = r->next;
			const uint64_t s0 = r->r0 + (uint64_t)bi * B, e0 = std::min<uint64_t>(r->r1, s0 + B);
			const int slot = (int)(seq % m->n_slots);
			trace_row(r, bi)[5] = host_ms();
			rc = cs_i_submit(d->ctx, slot, (uint32_t)(e0 - s0), r->j.off, r->j.bases, r->j.packed, r->j.nmask, s0, &r->j.opt);
			d->host_s[r->j.set][0] += since(t);
			if (rc != CS_OK) break;
			fl.push_back(Flight{r, bi, slot}); ++seq; ++r->next; progress = true;
		}
		if (rc != CS_OK) { fail_all(rc); continue; }
		// ---- kernels of the oldest unfinished batch done?  then its result copy goes out ----
		if (enq < fl.size()) {
			const auto t = now();
			Flight f = fl[enq];
			Run *r = f.run;
			BlockBuf &b = *r->b;
			const int pr = cs_i_poll(d->ctx, f.slot);
			if (pr < 0) { fail_all(pr); continue; }
			if (pr == 1) {
				uint64_t nm = 0, ns = 0;
				rc = cs_i_finish(d->ctx, f.slot, &nm, &ns);
				if (rc == CS_E_OVERFLOW) { // this batch needs larger slot buffers: drain, re-create the ctx once with what it needs, resubmit from here
					uint64_t need_m = 0, need_s = 0;
					cs_ctx_need(d->ctx, f.slot, &need_m, &need_s);
					rc = CS_OK;
					while (enq > 0 && rc == CS_OK) { const int c = complete_front(true); rc = c < 0 ? c : CS_OK; }
					if (rc != CS_OK) { fail_all(rc); continue; }
					if (need_m >= (1ull << 32) || need_s >= (1ull << 32)) {
						fail_all(cs_set_err(CS_E_OVERFLOW, "a batch of %u reads needs more than 2^32 mems or seeds: use smaller batches", m->batch_reads));
						continue;
					}
					if (need_m <= d->cap_mems && need_s <= d->cap_seeds) { need_m = d->cap_mems * 2; need_s = d->cap_seeds * 2; }
					d->cap_mems = std::max(d->cap_mems, need_m); d->cap_seeds = std::max(d->cap_seeds, need_s);
					for (const Flight &g : fl) g.run->next = std::min(g.run->next, g.bi);   // the batches still in flight are dropped with the old ctx
					fl.clear(); enq = 0;
					if ((rc = make_ctx(d)) != CS_OK) { fail_all(rc); continue; }
					{ float t5[5]; cs_i_slot_times(d->ctx, 0, t5, &base_ns); }
					continue;
				}
				if (rc != CS_OK) { fail_all(rc); continue; }
				{ // room for this batch in the block arrays (rare: the estimate per read was too low)
					const uint64_t mb = b.mem_base[f.bi], sb = b.seed_base[f.bi];
					const uint64_t cap1 = m->chaining ? b.cap_chains : b.cap_mems, cap2 = m->chaining ? std::min(b.cap_seeds, b.cap_sq) : b.cap_seeds;
					if (mb + nm > cap1 || sb + ns > cap2) {
						// copies into the old arrays must have landed (those of this set; an older set's go to its own arrays, but the order is kept)
						rc = CS_OK;
						while (enq > 0 && rc == CS_OK) { const int c = complete_front(true); rc = c < 0 ? c : CS_OK; }
						if (rc != CS_OK) { fail_all(rc); continue; }
						void *p; uint64_t c;
						if (m->chaining) {
							p = b.chains; c = b.cap_chains;
							if ((rc = pinned_grow(&p, &c, mb + nm, sizeof(cs_chain_t), mb)) == CS_OK) { b.chains = (cs_chain_t*)p; b.cap_chains = c; }
							p = b.sq; c = b.cap_sq;
							if (rc == CS_OK && (rc = pinned_grow(&p, &c, sb + ns, 2, sb)) == CS_OK) b.sq = (uint16_t*)p;
							p = b.sl; c = b.cap_sq;
							if (rc == CS_OK && (rc = pinned_grow(&p, &c, sb + ns, 2, sb)) == CS_OK) { b.sl = (uint16_t*)p; b.cap_sq = c; }
						} else {
							p = b.cmems; c = b.cap_mems;
							if ((rc = pinned_grow(&p, &c, mb + nm, sizeof(cs_cmem_t), mb)) == CS_OK) { b.cmems = (cs_cmem_t*)p; b.cap_mems = c; }
						}
						p = b.rlo; c = b.cap_seeds;
						if (rc == CS_OK && (rc = pinned_grow(&p, &c, sb + ns, 4, sb)) == CS_OK) b.rlo = (uint32_t*)p;
						p = b.rhi; c = b.cap_seeds;
						if (rc == CS_OK && (rc = pinned_grow(&p, &c, sb + ns, 1, sb)) == CS_OK) { b.rhi = (uint8_t*)p; b.cap_seeds = c; }
						if (rc != CS_OK) { fail_all(rc); continue; }
					}
					uint32_t *o1 = b.mem_off + (uint64_t)f.bi * (B + 1), *o2 = b.seed_off + (uint64_t)f.bi * (B + 1);
					if (m->chaining) rc = cs_i_fetch_chains_into(d->ctx, f.slot, o1, o2, b.chains + mb, b.rlo + sb, b.rhi + sb, b.sq + sb, b.sl + sb);
					else rc = cs_i_fetch_compact_into(d->ctx, f.slot, o1, o2, b.cmems + mb, b.rlo + sb, b.rhi + sb);
					if (rc != CS_OK) { fail_all(rc); continue; }
					b.mem_base[f.bi + 1] = mb + nm; b.seed_base[f.bi + 1] = sb + ns;
				}
				trace_row(r, f.bi)[6] = host_ms();
				++enq; ++r->done; progress = true;
			}
			d->host_s[r->j.set][1] += since(t);
		}
		// ---- results of the oldest batch on the host? ----
		if (enq > 0) {
			const auto t = now();
			Run *r = fl.front().run;
			const int c = complete_front(false);
			if (c < 0) { fail_all(c); continue; }
			if (c == 1) progress = true;
			d->host_s[r->j.set][1] += since(t);
		}
		if (!progress) { const auto t = now(); std::this_thread::sleep_for(std::chrono::microseconds(20)); d->host_s[runs.front()->j.set][2] += since(t); }
	}
}

} // namespace

extern "C" cs_multi_t *cs_multi_create(cs_index_t *const *idx, int n_dev, uint32_t batch_reads, uint32_t max_read_len, int n_slots,
                                       uint32_t mems_per_read, uint32_t seeds_per_read, const cs_ctx_config_t *cfg)
{
	if (!idx || n_dev < 1 || n_dev > CS_MULTI_MAX_DEV || batch_reads == 0 || max_read_len == 0 || n_slots < 2 || n_slots > 16) {
		cs_set_err(CS_E_ARG, "bad multi geometry (%d devices (1..%d), %u reads per batch, %d slots (2..16))", n_dev, CS_MULTI_MAX_DEV, batch_reads, n_slots);
		return nullptr;
	}
	if ((uint64_t)batch_reads * max_read_len >= (1ull << 32)) { cs_set_err(CS_E_ARG, "batch_reads x max_read_len must stay below 2^32 bases"); return nullptr; }
	cs_multi *m = new cs_multi();
	m->n_dev = n_dev; m->batch_reads = batch_reads; m->max_read_len = max_read_len; m->n_slots = n_slots;
	m->mems_per_read = mems_per_read ? mems_per_read : 16; m->seeds_per_read = seeds_per_read ? seeds_per_read : 32;
	if (cfg) m->cfg = *cfg; else cs_ctx_config_default(&m->cfg);
	m->busy[0] = m->busy[1] = false; m->chaining = false; m->l_pac = 0;
	for (int k = 0; k < n_dev; ++k) {
		Dev *d = new Dev();
		d->m = m; d->k = k; d->idx = idx[k]; d->ctx = nullptr;
		d->cap_mems = (uint64_t)batch_reads * m->mems_per_read; d->cap_seeds = (uint64_t)batch_reads * m->seeds_per_read;
		d->done[0] = d->done[1] = false; d->rc[0] = d->rc[1] = CS_OK;
		m->dev.push_back(d);
		if (!idx[k] || make_ctx(d) != CS_OK) { if (!idx[k]) cs_set_err(CS_E_ARG, "null index for device slot %d", k); cs_multi_free(m); return nullptr; }
	}
	for (Dev *d : m->dev) d->th = std::thread(worker, d);
	return m;
}

extern "C" void cs_multi_free(cs_multi_t *m)
{
	if (!m) return;
	for (Dev *d : m->dev) {
		if (d->th.joinable()) {
			{ std::lock_guard<std::mutex> lk(d->mu); d->stop = true; }
			d->cv.notify_all();
			d->th.join();
		}
		if (d->ctx) cs_ctx_free(d->ctx);
		if (d->idx) cudaSetDevice(d->idx->device);
		for (int s = 0; s < 2; ++s) {
			BlockBuf &b = d->buf[s];
			cudaFreeHost(b.mem_base); cudaFreeHost(b.seed_base); cudaFreeHost(b.mem_off); cudaFreeHost(b.seed_off);
			cudaFreeHost(b.cmems); cudaFreeHost(b.rlo); cudaFreeHost(b.rhi); cudaFreeHost(b.chains); cudaFreeHost(b.sq); cudaFreeHost(b.sl);
		}
		delete d;
	}
	delete m;
}

static int multi_submit(cs_multi_t *m, int set, uint64_t n_reads, const uint8_t *bases, const uint64_t *packed, const uint32_t *nmask,
                        const uint64_t *offsets, const cs_seed_opt_t *opt)
{
	if (!m || set < 0 || set > 1 || !offsets || !opt || (!bases && !(packed && nmask))) return cs_set_err(CS_E_ARG, "bad argument");
	if (m->busy[set]) return cs_set_err(CS_E_STATE, "read set %d is still in flight: cs_multi_wait it first", set);
	if (offsets[0] != 0) return cs_set_err(CS_E_ARG, "offsets[0] must be 0");
	Job j;
	j.set = set; j.n_reads = n_reads; j.bases = bases; j.packed = packed; j.nmask = nmask; j.off = offsets; j.opt = *opt;
	m->busy[set] = true; m->n_reads[set] = n_reads; m->t0[set] = std::chrono::steady_clock::now();
	for (Dev *d : m->dev) {
		{ std::lock_guard<std::mutex> lk(d->mu); d->done[set] = false; d->q.push_back(j); }
		d->cv.notify_all();
	}
	return CS_OK;
}

extern "C" int cs_multi_submit(cs_multi_t *m, int set, uint64_t n_reads, const uint8_t *bases, const uint64_t *offsets, const cs_seed_opt_t *opt)
{ return multi_submit(m, set, n_reads, bases, nullptr, nullptr, offsets, opt); }

extern "C" int cs_multi_submit_packed(cs_multi_t *m, int set, uint64_t n_reads, const uint64_t *packed, const uint32_t *nmask,
                                      const uint64_t *offsets, const cs_seed_opt_t *opt)
{ return multi_submit(m, set, n_reads, nullptr, packed, nmask, offsets, opt); }

extern "C" int cs_multi_wait(cs_multi_t *m, int set, cs_multi_result_t *out)
{
	if (!m || set < 0 || set > 1 || !out) return cs_set_err(CS_E_ARG, "bad argument");
	if (!m->busy[set]) return cs_set_err(CS_E_STATE, "read set %d was not submitted", set);
	int rc = CS_OK;
	memset(out, 0, sizeof *out);
	for (Dev *d : m->dev) {
		std::unique_lock<std::mutex> lk(d->mu);
		d->cv.wait(lk, [&] { return d->done[set]; });
		if (d->rc[set] != CS_OK && rc == CS_OK) { rc = d->rc[set]; cs_set_err(rc, "device %d: %s", d->idx->device, d->err[set]); }
		m->blocks[set][d->k] = d->buf[set].pub;
		out->n_mems += d->n_mems[set]; out->n_seeds += d->n_seeds[set];
		out->counters.ext_queries += d->cnt[set].ext_queries; out->counters.ext_calls += d->cnt[set].ext_calls;
		out->counters.sal_queries += d->cnt[set].sal_queries; out->counters.sal_calls += d->cnt[set].sal_calls;
		if (d->host_s[set][0] + d->host_s[set][1] + d->host_s[set][2] > out->host_s[0] + out->host_s[1] + out->host_s[2])
			for (int q = 0; q < 3; ++q) out->host_s[q] = d->host_s[set][q];
		for (int q = 0; q < 4; ++q) out->gpu_ms[q] = std::max(out->gpu_ms[q], d->gpu_ms[set][q]);
	}
	m->busy[set] = false;
	out->n_reads = m->n_reads[set]; out->n_blocks = m->n_dev; out->blocks = m->blocks[set];
	out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - m->t0[set]).count();
	return rc;
}

extern "C" int cs_multi_set_chaining(cs_multi_t *m, const cs_bns_view_t *bns, const cs_chain_opt_t *opt)
{
	if (!m) return cs_set_err(CS_E_ARG, "null argument");
	if (m->busy[0] || m->busy[1]) return cs_set_err(CS_E_STATE, "a read set is in flight");
	if (bns) {
		if (!opt || bns->n_seqs < 1 || !bns->offset) return cs_set_err(CS_E_ARG, "bad contig table");
		m->l_pac = bns->l_pac; m->copt = *opt;
		m->c_off.assign(bns->offset, bns->offset + bns->n_seqs);
		if (bns->is_alt) m->c_alt.assign(bns->is_alt, bns->is_alt + bns->n_seqs); else m->c_alt.clear();
	}
	m->chaining = bns != nullptr;
	for (Dev *d : m->dev) {
		std::lock_guard<std::mutex> lk(d->mu);     // (the worker is idle: no set in flight)
		cs_bns_view_t v; v.l_pac = m->l_pac; v.n_seqs = (int32_t)m->c_off.size(); v.offset = m->c_off.data(); v.is_alt = m->c_alt.empty() ? nullptr : m->c_alt.data();
		const int rc = cs_ctx_set_chaining(d->ctx, m->chaining ? &v : nullptr, &m->copt);
		if (rc != CS_OK) { m->chaining = false; return rc; }
	}
	return CS_OK;
}

extern "C" void cs_multi_block_bounds(uint64_t n_reads, int n_dev, int k, uint64_t *r0, uint64_t *r1)
{
	if (n_dev < 1 || k < 0 || k >= n_dev || !r0 || !r1) { if (r0) *r0 = 0; if (r1) *r1 = 0; return; }
	block_bounds(n_reads, n_dev, k, r0, r1);
}

extern "C" uint32_t cs_multi_trace(const cs_multi_t *m, int set, int k, float *out, uint32_t cap_batches)
{ // diagnostics: the timeline of the batches device k ran for the last finished run of `set` (see compseed_b200.h)
	if (!m || set < 0 || set > 1 || k < 0 || k >= m->n_dev) return 0;
	const std::vector<float> &t = m->dev[k]->trace[set];
	const uint32_t nb = (uint32_t)(t.size() / CS_TRACE_COLS);
	if (out) memcpy(out, t.data(), (size_t)std::min(nb, cap_batches) * CS_TRACE_COLS * sizeof(float));
	return nb;
}

extern "C" uint64_t cs_multi_launches(const cs_multi_t *m)
{
	uint64_t n = 0;
	if (m) for (Dev *d : m->dev) n += cs_ctx_launches(d->ctx);
	return n;
}

extern "C" int cs_multi_gather(const cs_multi_result_t *res, uint64_t *mem_off, cs_mem_t *mems, uint64_t *seed_off, int64_t *rbeg, int n_threads)
{ // flat arrays in input order (tests, hosts that want them); plain host C++
	if (!res || !mem_off || !seed_off) return cs_set_err(CS_E_ARG, "null argument");
	if (res->n_blocks > 0 && res->blocks[0].chains) return cs_set_err(CS_E_STATE, "this result holds chains (cs_multi_read_chains), not mems");
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	uint64_t mb = 0, sb = 0;
	mem_off[0] = seed_off[0] = 0;
	for (int k = 0; k < res->n_blocks; ++k) {
		const cs_block_t *b = &res->blocks[k];
		const uint64_t n = b->r1 - b->r0;
		if (n == 0) continue;
		const uint64_t nm = b->mem_base[b->n_batches], ns = b->seed_base[b->n_batches];
		auto work = [=](int t) {
			for (uint64_t r = n * t / n_threads; r < n * (t + 1) / n_threads; ++r) {
				const uint64_t bi = r / b->batch_reads, lr = r % b->batch_reads;
				mem_off[b->r0 + r + 1] = mb + b->mem_base[bi] + b->mem_off[bi * (b->batch_reads + 1ull) + lr + 1];
				seed_off[b->r0 + r + 1] = sb + b->seed_base[bi] + b->seed_off[bi * (b->batch_reads + 1ull) + lr + 1];
			}
			if (mems) for (uint64_t i = nm * t / n_threads; i < nm * (t + 1) / n_threads; ++i) cs_cmem_unpack(b->cmems + i, mems + mb + i);
			if (rbeg) for (uint64_t i = ns * t / n_threads; i < ns * (t + 1) / n_threads; ++i) rbeg[sb + i] = cs_crbeg(b->rbeg_lo, b->rbeg_hi, i);
		};
		std::vector<std::thread> th;
		for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
		for (auto &t : th) t.join();
		mb += nm; sb += ns;
	}
	return CS_OK;
}

extern "C" int cs_pack_reads_host64(uint64_t n_reads, const uint8_t *bases, const uint64_t *offsets, uint64_t *packed, uint32_t *nmask, int n_threads)
{ // the set-global packed layout: read r owns the words [(offsets[r] >> 5) + 2r, ... + (len_r >> 5) + 2); everything else is "all N"
	if (!bases || !offsets || !packed || !nmask) return cs_set_err(CS_E_ARG, "null argument");
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	const uint64_t nw_tot = (offsets[n_reads] >> 5) + 2 * n_reads;
	auto work = [=](int t) {
		for (uint64_t r = n_reads * t / n_threads; r < n_reads * (t + 1) / n_threads; ++r) {
			const uint64_t o = offsets[r], len = offsets[r + 1] - o;
			const uint64_t w0 = (o >> 5) + 2 * r, nw = (len >> 5) + 2;
			const uint8_t *q = bases + o;
			for (uint64_t w = 0; w < nw; ++w) {
				uint64_t v = 0; uint32_t mk = 0xffffffffu;
				const uint64_t p0 = w << 5, cnt = p0 < len ? std::min<uint64_t>(len - p0, 32) : 0;
				for (uint64_t jj = 0; jj < cnt; ++jj) {
					const uint32_t c = q[p0 + jj];
					if (c <= 3) { v |= (uint64_t)c << (2 * jj); mk &= ~(1u << jj); }
				}
				packed[w0 + w] = v; nmask[w0 + w] = mk;
			}
			const uint64_t next0 = r + 1 < n_reads ? (offsets[r + 1] >> 5) + 2 * (r + 1) : nw_tot;
			for (uint64_t w = w0 + nw; w < next0; ++w) { packed[w] = 0; nmask[w] = 0xffffffffu; }
		}
	};
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
	for (auto &t : th) t.join();
	return CS_OK;
}

// Replica of an index on another device: device-to-device copies (over NVLink between peers), no host staging.
extern "C" cs_index_t *cs_index_replicate(const cs_index_t *src, int device)
{
	if (!src) { cs_set_err(CS_E_ARG, "null index"); return nullptr; }
	if (cs_use_device(device) != CS_OK) return nullptr;
	cs_index *idx = (cs_index*)calloc(1, sizeof(cs_index));
	if (!idx) { cs_set_err(CS_E_ARG, "out of host memory"); return nullptr; }
	*idx = *src;
	idx->device = device;
	idx->d_buckets = nullptr; idx->d_sa = nullptr; idx->d_kt = nullptr; idx->d_pt = nullptr; idx->d_text = nullptr; idx->d_isa = nullptr; idx->d_rep = nullptr;
	{
		int can = 0;
		if (device != src->device && cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
			cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
			if (e != cudaSuccess) cudaGetLastError();   // already enabled is fine; the copies below work either way
		}
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		idx->n_sm = prop.multiProcessorCount;
		const DevIndex &S = src->d;
		auto clone = [&](void **dst, const void *from, uint64_t bytes) -> cudaError_t {
			if (!from || !bytes) { *dst = nullptr; return cudaSuccess; }
			cudaError_t e = cudaMalloc(dst, bytes);
			if (e != cudaSuccess) return e;
			return cudaMemcpyPeer(*dst, device, from, src->device, bytes);
		};
		CK(clone((void**)&idx->d_buckets, S.buckets, S.n_buckets * 32));
		CK(clone((void**)&idx->d_sa, S.sa, S.n_sa * 8));
		CK(clone((void**)&idx->d_kt, S.kt, S.kt ? (((1ull << (2 * (S.kt_depth + 1))) - 4) / 3) * 16 : 0));
		CK(clone((void**)&idx->d_pt, S.pt, S.pt ? ((1ull << (2 * S.pt_k)) / 16) * 4 : 0));
		CK(clone((void**)&idx->d_text, S.text, S.text ? ((S.seq_len + 31) / 32 + 2) * 8 : 0));
		CK(clone((void**)&idx->d_isa, S.isa, S.isa ? ((S.seq_len >> S.isa_shift) + 2) * 8 : 0));
		CK(clone((void**)&idx->d_rep, S.rep, S.rep ? ((S.seq_len + 63) & ~63ull) + 64 : 0));
		CK(cudaDeviceSynchronize());
		idx->d.rep = idx->d_rep;
		idx->d.buckets = idx->d_buckets; idx->d.sa = idx->d_sa; idx->d.kt = idx->d_kt; idx->d.pt = idx->d_pt; idx->d.text = idx->d_text; idx->d.isa = idx->d_isa;
	}
	return idx;
fail:
	cudaFree(idx->d_buckets); cudaFree(idx->d_sa); cudaFree(idx->d_kt); cudaFree(idx->d_pt); cudaFree(idx->d_text); cudaFree(idx->d_isa); cudaFree(idx->d_rep);
	free(idx);
	return nullptr;
}
