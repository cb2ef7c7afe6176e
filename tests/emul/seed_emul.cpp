// TEST INFRASTRUCTURE ONLY.  Compiles compseed_b200/csrc/cs_kernels.cu as plain C++ (CUDA qualifiers defined away) and runs its kernels
// on the CPU: the index construction kernels (re-layout, dense SA, top-of-search table, 2-bit text, occurrence filter, inverse SA,
// repeat lengths) serially, and the seeding kernels in one of two ways --
//   seed_emul_run:   warps of ONE lane (a vote is the lane's own predicate, a shuffle returns the lane's own value), one host thread:
//                    k_seed_fast, k_seed_walk, k_seed_r3_fast; the calls handed on to the literal kernel are resolved with the oracle's
//                    bwt_smem1a restatement, exactly as k_seed's call mode defines them;
//   seed_emul_run32: REAL warps, one host thread per lane, every warp intrinsic a rendezvous of the 32: all seeding kernels, the literal
//                    k_seed (call mode and read mode) and the general third pass k_seed_r3 included; nothing is resolved by the oracle.
// tests/test_seed_emul.py compares the outcome with the oracle.  The shipped library never contains or runs this build: there is no
// CPU path in the product.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <thread>
#include <pthread.h>

#define CS_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__
struct emul_dim3 { unsigned x, y, z; };
static thread_local emul_dim3 threadIdx = {0, 0, 0};
static emul_dim3 blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T, class U> static inline T atomicAdd(T *p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
static inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicCAS(uint32_t *p, uint32_t cmp, uint32_t v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline uint32_t atomicExch(uint32_t *p, uint32_t v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline void __syncthreads() {}
// Warp intrinsics.  g_lanes == 1: a warp of ONE lane (a vote is the lane's own predicate, a shuffle returns its own value) -- the
// kernels then run in a single host thread.  g_lanes == 32: a real warp, one host thread per lane (emul_warp below); every intrinsic is a
// rendezvous of the 32 threads (all masks in the seeding kernels are full), so the vote-controlled loops, the warp-aggregated
// atomics and the cooperative occurrence filter of the literal kernel execute with their real semantics.
static int g_lanes = 1;
static pthread_barrier_t g_bar;
static uint64_t g_slot[32];
static inline void warp_rendezvous() { pthread_barrier_wait(&g_bar); }
static inline unsigned __ballot_sync(unsigned, bool p)
{
	if (g_lanes == 1) return p ? 1u : 0u;
	g_slot[threadIdx.x & 31] = p ? 1u : 0u;
	warp_rendezvous();
	unsigned m = 0;
	for (int i = 0; i < 32; ++i) m |= (unsigned)(g_slot[i] & 1u) << i;
	warp_rendezvous();
	return m;
}
static inline bool __any_sync(unsigned mk, bool p) { return __ballot_sync(mk, p) != 0; }
static inline bool __all_sync(unsigned mk, bool p) { return g_lanes == 1 ? p : __ballot_sync(mk, p) == 0xffffffffu; }
template <class T> static inline T emul_exchange(T v, int src)
{
	static_assert(sizeof(T) <= 8, "shuffle of more than 8 bytes");
	uint64_t raw = 0; memcpy(&raw, &v, sizeof v);
	g_slot[threadIdx.x & 31] = raw;
	warp_rendezvous();
	raw = g_slot[src & 31];
	warp_rendezvous();
	T r; memcpy(&r, &raw, sizeof r);
	return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{ if (g_lanes == 1) return v; const int lane = threadIdx.x & 31; return emul_exchange(v, (lane & ~(width - 1)) | (src & (width - 1))); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32)
{ if (g_lanes == 1) return v; const int lane = threadIdx.x & 31; const int src = (lane & (width - 1)) >= d ? lane - d : lane; return emul_exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int d, int width = 32)
{ if (g_lanes == 1) return v; const int lane = threadIdx.x & 31; return emul_exchange(v, lane ^ d); }
static inline void __syncwarp(unsigned = 0xffffffffu) { if (g_lanes != 1) warp_rendezvous(); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline uint32_t __brev(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline uint64_t __brevll(uint64_t v) { uint64_t r = 0; for (int i = 0; i < 64; ++i) r |= ((v >> i) & 1ull) << (63 - i); return r; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t __vcmpgeu4(uint32_t a, uint32_t b)
{ uint32_t r = 0; for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xff) >= ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline uint32_t __vcmpgtu4(uint32_t a, uint32_t b)
{ uint32_t r = 0; for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xff) > ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
using std::min; using std::max;

uint4 s_dyn[1 << 16], s_list[1 << 16];   // the kernels' dynamic shared memory (extern __shared__ in their bodies)

#include "../../compseed_b200/csrc/cs_kernels.cu"
#include "../../oracle/cs_oracle.h"

namespace {
struct Emul {
	DevIndex d;
	cso_index_t oidx;
	std::vector<uint4> buckets, kt;
	std::vector<uint64_t> sa, text, isa;
	std::vector<unsigned long long> W;
	std::vector<uint32_t> pt;
	std::vector<uint8_t> rep;
};
void one_thread() { blockDim.x = 1; gridDim.x = 1; threadIdx.x = 0; blockIdx.x = 0; }
}

// one CTA of one real warp: 32 host threads, lane = threadIdx.x
template <class K> static void emul_warp(K kernel)
{
	pthread_barrier_init(&g_bar, nullptr, 32);
	g_lanes = 32;
	blockDim.x = 32; gridDim.x = 1; blockIdx.x = 0;
	std::vector<std::thread> th;
	for (unsigned lane = 0; lane < 32; ++lane) th.emplace_back([=]() { threadIdx.x = lane; kernel(); });
	for (auto &t : th) t.join();
	g_lanes = 1;
	pthread_barrier_destroy(&g_bar);
	one_thread();
}

extern "C" {

// Builds the device index of cs_index_upload_ex (dense SA) on the CPU with the library's own kernels.  K, depth, isa_intv: as
// cs_index_config_t (must be given; the defaults depend on seq_len).  use_rep: also the repeat-length array.
void *seed_emul_index(uint64_t primary, const uint64_t *L2, uint64_t seq_len, const uint32_t *bwt, uint64_t bwt_size, const uint64_t *sa_in,
                      uint64_t n_sa_in, int sa_intv, int K, int depth, int isa_intv, int use_rep)
{
	Emul *E = new Emul();
	DevIndex &d = E->d;
	memset(&d, 0, sizeof d);
	one_thread();
	E->oidx.primary = primary; memcpy(E->oidx.L2, L2, 40); E->oidx.seq_len = seq_len; E->oidx.bwt_size = bwt_size; E->oidx.bwt = const_cast<uint32_t*>(bwt);
	E->oidx.sa_intv = sa_intv; E->oidx.n_sa = n_sa_in; E->oidx.sa = const_cast<uint64_t*>(sa_in);
	d.primary = primary; d.seq_len = seq_len; memcpy(d.L2, L2, 40);
	d.n_buckets = (seq_len + 63) / 64 + 1;
	E->buckets.assign(d.n_buckets * 2, make_uint4(0, 0, 0, 0));
	k_relayout(bwt, bwt_size, seq_len, E->buckets.data(), d.n_buckets - 1);
	d.buckets = E->buckets.data();
	std::vector<uint64_t> sa0(sa_in, sa_in + n_sa_in);
	sa0[0] = (uint64_t)-1;
	int sh = 0; while ((1 << sh) < sa_intv) ++sh;
	d.sa = sa0.data(); d.n_sa = n_sa_in; d.sa_mask = (uint32_t)sa_intv - 1; d.sa_shift = (uint32_t)sh;
	E->sa.assign(seq_len + 1, 0);
	k_resample_sa(d, E->sa.data(), seq_len + 1, 0);
	d.sa = E->sa.data(); d.n_sa = seq_len + 1; d.sa_mask = 0; d.sa_shift = 0;
	// top-of-search table
	E->kt.assign(((1ull << (2 * (depth + 1))) - 4) / 3, make_uint4(0, 0, 0, 0));
	for (int dd = 1; dd <= depth; ++dd) k_kt_build(d, E->kt.data(), (uint32_t)dd);
	d.kt = E->kt.data(); d.kt_depth = (uint32_t)depth;
	// text, filter, inverse SA
	const uint64_t n_words = (seq_len + 31) / 32 + 2;
	E->W.assign(n_words, 0);
	k_text_from_index(d, E->W.data());
	E->pt.assign((1ull << (2 * K)) / 16 + 1, 0);
	k_pt_count(reinterpret_cast<const uint64_t*>(E->W.data()), seq_len, (uint32_t)K, E->pt.data());
	d.pt = E->pt.data(); d.pt_k = (uint32_t)K;
	E->text.assign(n_words, 0);
	k_text_lsb(reinterpret_cast<const uint64_t*>(E->W.data()), n_words, E->text.data());
	int ish = 0; while ((1 << ish) < isa_intv) ++ish;
	E->isa.assign((seq_len >> ish) + 2, 0);
	k_isa_sample(d, E->isa.data(), (uint32_t)ish);
	d.text = E->text.data(); d.isa = E->isa.data(); d.isa_shift = (uint32_t)ish;
#ifdef CS_HAVE_REP
	if (use_rep) {
		E->rep.assign(seq_len + 64, 0);
		k_rep_build(d, E->rep.data());
		d.rep = E->rep.data();
	}
#endif
	return E;
}

void seed_emul_free(void *e) { delete static_cast<Emul*>(e); }
const uint8_t *seed_emul_rep(void *e) { Emul *E = static_cast<Emul*>(e); return E->rep.empty() ? nullptr : E->rep.data(); }
const uint64_t *seed_emul_text(void *e) { return static_cast<Emul*>(e)->text.data(); }
const uint64_t *seed_emul_sa(void *e) { return static_cast<Emul*>(e)->sa.data(); }

// Passes 1 and 2 of mem_collect_intv for a batch: k_seed_fast + k_seed_walk emulated, the rest resolved by the oracle.
// out_mems[cap] per read sorted by info, mem_off[n+1].  stats[0..3]: counters (ext queries, FM extends, two-sector, filter probes),
// [4..5] executed requests fast / walk, [6] deferred calls, [7] of them for the literal kernel, [8] resolved by the oracle's
// literal call, [16..31] the CS_STATS event counters of k_seed_fast.  Returns the number of mems, or < 0 (error code / -100: cap).
int64_t seed_emul_run(void *e, uint32_t n_reads, const uint8_t *bases, const uint32_t *off, const cs_seed_opt_t *opt,
                      cs_mem_t *out_mems, uint64_t cap, uint32_t *mem_off, uint64_t *stats)
{
	Emul *E = static_cast<Emul*>(e);
	const DevIndex &d = E->d;
	const uint64_t n_bases = off[n_reads];
	uint32_t max_len = 0;
	for (uint32_t r = 0; r < n_reads; ++r) max_len = std::max(max_len, off[r + 1] - off[r]);
	if (max_len + 32 > 32 * CS_READ_SMEM) return -101;
	std::vector<uint8_t> pb(bases, bases + n_bases); pb.resize(n_bases + 256, 0);
	std::vector<uint64_t> packed((n_bases >> 5) + 2 * (size_t)n_reads + 8, 0);
	std::vector<uint32_t> nmask(packed.size(), 0);
	blockDim.x = 8; gridDim.x = 1; blockIdx.x = 0;
	for (unsigned t = 0; t < 8; ++t) { threadIdx.x = t; k_pack_reads(pb.data(), off, n_reads, packed.data(), nmask.data()); }
	one_thread();
	const uint32_t mem_cap = std::min<uint32_t>(2 * max_len + 16, 4096), defer_cap = 8 * n_reads + 4096;
	std::vector<uint32_t> ctrl(64, 0);
	std::vector<unsigned long long> counters(64, 0), req(8, 0);
	unsigned long long pool_used = 0;
	int error = 0;
	uint32_t n_defer = 0, n_lit = 0, n_defer_fast = 0;
	std::vector<uint4> defer_q(defer_cap), defer_lx(defer_cap);
	uint4 thread_lx[1];
	std::vector<uint32_t> defer_bits(defer_cap, 0), lit_q(defer_cap, 0), x_n(defer_cap, 0xffffffffu), read_last_q(n_reads + 1, 0xffffffffu), read_n_mems(n_reads + 1, 0);
	std::vector<uint64_t> x_off(defer_cap, 0), read_pool_off(n_reads + 1, 0);
	std::vector<cs_mem_t> thread_mems(mem_cap), pool(cap);
	SeedArgs a;
	memset(&a, 0, sizeof a);
	a.bases = pb.data(); a.off = off; a.n_reads = n_reads; a.opt = *opt; a.packed = packed.data(); a.off_bias = 0; a.nmask = nmask.data();
	a.next_read = ctrl.data(); a.defer_q = defer_q.data(); a.defer_bits = defer_bits.data(); a.defer_lx = defer_lx.data(); a.thread_lx = thread_lx; a.lit_q = lit_q.data(); a.n_lit = &n_lit;
	a.defer_cap = defer_cap; a.n_defer = &n_defer; a.n_defer_fast = &n_defer_fast; a.read_last_q = read_last_q.data();
	a.x_off = x_off.data(); a.x_n = x_n.data(); a.thread_mems = thread_mems.data(); a.mem_cap = mem_cap;
	a.pool = pool.data(); a.pool_cap = cap; a.pool_used = &pool_used; a.read_pool_off = read_pool_off.data(); a.read_n_mems = read_n_mems.data();
	a.counters = counters.data(); a.req = req.data(); a.error = &error;
	k_seed_fast(d, a);
	n_defer_fast = n_defer;
	const uint32_t n_lit_fast = n_lit;
	std::vector<uint32_t> walk_order(defer_cap, 0xffffffffu);
	uint32_t hist[64] = {0}, cursor[64] = {0}, n_walk = 0;
	a.walk_order = walk_order.data(); a.n_walk = &n_walk;
	// (k_walk_count / k_walk_scatter use CTA-wide shared histograms and barriers: not emulated; the same counting sort with the kernels' cost function)
	for (uint32_t qq = 0; qq < n_defer_fast; ++qq) if (defer_q[qq].y >> 31) ++hist[walk_cost_class(defer_q[qq], defer_bits[qq], (int)d.pt_k, (int)d.kt_depth)];
	k_walk_scan(hist, cursor, &n_walk);
	for (uint32_t qq = 0; qq < n_defer_fast; ++qq) if (defer_q[qq].y >> 31) walk_order[cursor[walk_cost_class(defer_q[qq], defer_bits[qq], (int)d.pt_k, (int)d.kt_depth)]++] = qq;
	if (n_defer > defer_cap) return -102;
	k_seed_walk(d, a);
	if (error) return error;
	// third pass (text-assisted kernel; reads the first-pass SMEMs k_seed_fast left in the pool), when the options ask for it
	const uint32_t kp1 = (uint32_t)opt->min_seed_len + 1;
	std::vector<cs_mem_t> r3_mems((size_t)n_bases / kp1 + n_reads + 2);
	std::vector<uint32_t> r3_n(n_reads + 1, 0);
	if (opt->max_mem_intv > 0) {
		a.r3_mems = r3_mems.data(); a.r3_n_mems = r3_n.data();
		k_seed_r3_fast(d, a);
		if (error) return error;
	}
	// per read: what the kernels stored ...
	std::vector<std::vector<cs_mem_t>> per(n_reads);
	if (opt->max_mem_intv > 0)
		for (uint32_t r = 0; r < n_reads; ++r)
			for (uint32_t m = 0; m < r3_n[r]; ++m) per[r].push_back(r3_mems[(size_t)(off[r] / kp1) + r + m]);
	for (uint32_t r = 0; r < n_reads; ++r)
		for (uint32_t m = 0; m < read_n_mems[r]; ++m) per[r].push_back(pool[read_pool_off[r] + m]);
	for (uint32_t q = 0; q < n_defer_fast; ++q)
		if ((defer_q[q].y >> 31) && x_n[q] != 0xffffffffu)
			for (uint32_t m = 0; m < x_n[q]; ++m) per[defer_q[q].x].push_back(pool[x_off[q] + m]);
	// ... and the calls listed for the literal kernel, executed as k_seed's call mode defines them: the whole bwt_smem1a call
	// (mems of >= min_seed_len bases), and for a first-pass call the second-pass calls of what it found (bwamem.c:238-249)
	uint64_t n_oracle = 0, n_punt = 0, n_follow = 0;
	std::vector<cso_mem_t> tmp(4096), tmp2(4096);
	for (uint32_t li = 0; li < n_lit; ++li) {
		const uint4 it = defer_q[lit_q[li]];
		if (li >= n_lit_fast) { if (lit_q[li] < n_defer_fast) ++n_punt; else ++n_follow; }
		const uint32_t rd = it.x; const int pivot = (int)(it.y & 0xffff), pass = (int)((it.y >> 16) & 3);
		const uint8_t *q = bases + off[rd]; const int len = (int)(off[rd + 1] - off[rd]);
		int n1 = cso_smem1_call(&E->oidx, len, q, pivot, it.z, tmp.data(), (int)tmp.size(), nullptr);
		if (n1 < 0) return -103;
		++n_oracle;
		for (int i = 0; i < n1; ++i) {
			const int s = (int)(tmp[i].info >> 32), en = (int)(uint32_t)tmp[i].info;
			if (en - s < opt->min_seed_len) continue;
			cs_mem_t m; memcpy(&m, &tmp[i], sizeof m);
			per[rd].push_back(m);
			if (pass != 1 || en - s < opt->split_len || tmp[i].x[2] > (uint64_t)opt->split_width) continue;
			int n2 = cso_smem1_call(&E->oidx, len, q, (s + en) >> 1, tmp[i].x[2] + 1, tmp2.data(), (int)tmp2.size(), nullptr);
			if (n2 < 0) return -103;
			for (int j = 0; j < n2; ++j)
				if ((int)((uint32_t)tmp2[j].info - (uint32_t)(tmp2[j].info >> 32)) >= opt->min_seed_len) { cs_mem_t m2; memcpy(&m2, &tmp2[j], sizeof m2); per[rd].push_back(m2); }
		}
	}
	uint64_t n = 0;
	mem_off[0] = 0;
	for (uint32_t r = 0; r < n_reads; ++r) {
		std::stable_sort(per[r].begin(), per[r].end(), [](const cs_mem_t &x, const cs_mem_t &y) { return x.info < y.info; });
		if (n + per[r].size() > cap) return -100;
		for (const cs_mem_t &m : per[r]) out_mems[n++] = m;
		mem_off[r + 1] = (uint32_t)n;
	}
	if (stats) {
		for (int k = 0; k < 4; ++k) stats[k] = counters[k];
		stats[12] = req[3];   // executed requests of the third-pass kernel
		stats[4] = req[0]; stats[5] = req[1]; stats[6] = n_defer; stats[7] = n_lit; stats[8] = n_oracle;
		for (int k = 40; k < 48; ++k) stats[k] = counters[k];
		stats[9] = n_lit_fast; stats[10] = n_punt; stats[11] = n_follow;   // literal tasks: straight from k_seed_fast, punted by k_seed_walk, second-pass follow-ups of what k_seed_walk found
		for (int k = 0; k < 16; ++k) stats[16 + k] = counters[20 + k];
	}
	return (int64_t)n;
}

// The same batch with REAL warps (32 host threads per warp, see emul_warp) and every seeding kernel of the library, the literal one
// included; nothing is resolved by the oracle.  mode 0: k_seed_fast -> k_seed_walk -> k_seed in call mode -> k_seed_r3_fast (what a
// batch runs with the dense SA); mode 1: k_seed alone in read mode -> k_seed_r3 (what it runs without the fast kernels).
// Returns the number of mems (all passes the options ask for), or < 0.
int64_t seed_emul_run32(void *e, uint32_t n_reads, const uint8_t *bases, const uint32_t *off, const cs_seed_opt_t *opt, int mode,
                        cs_mem_t *out_mems, uint64_t cap, uint32_t *mem_off, uint64_t *stats)
{
	Emul *E = static_cast<Emul*>(e);
	const DevIndex &d = E->d;
	const uint64_t n_bases = off[n_reads];
	uint32_t max_len = 0;
	for (uint32_t r = 0; r < n_reads; ++r) max_len = std::max(max_len, off[r + 1] - off[r]);
	if (max_len + 32 > 32 * CS_READ_SMEM) return -101;
	std::vector<uint8_t> pb(bases, bases + n_bases); pb.resize(n_bases + 256, 0);
	std::vector<uint64_t> packed((n_bases >> 5) + 2 * (size_t)n_reads + 8, 0);
	std::vector<uint32_t> nmask(packed.size(), 0);
	blockDim.x = 8; gridDim.x = 1; blockIdx.x = 0;
	for (unsigned t = 0; t < 8; ++t) { threadIdx.x = t; k_pack_reads(pb.data(), off, n_reads, packed.data(), nmask.data()); }
	one_thread();
	const uint32_t mem_cap = std::min<uint32_t>(2 * max_len + 16, 4096), defer_cap = 8 * n_reads + 4096;
	const uint32_t spill_cap = max_len > CS_LIST_SMEM ? max_len - CS_LIST_SMEM + 1 : 1;
	const size_t nthreads = CS_SEED_BLOCK;                       // one CTA; its first warp runs
	std::vector<uint32_t> ctrl(64, 0);
	std::vector<unsigned long long> counters(64, 0), req(8, 0);
	unsigned long long pool_used = 0;
	int error = 0;
	uint32_t n_defer = 0, n_lit = 0, n_defer_fast = 0, n_walk = 0;
	std::vector<uint4> defer_q(defer_cap), defer_lx(defer_cap), thread_lx(nthreads), spill((size_t)spill_cap * nthreads);
	std::vector<uint32_t> defer_bits(defer_cap, 0), lit_q(defer_cap, 0), x_n(defer_cap, 0xffffffffu), read_last_q(n_reads + 1, 0xffffffffu), read_n_mems(n_reads + 1, 0);
	std::vector<uint32_t> walk_order(defer_cap, 0xffffffffu);
	std::vector<uint64_t> x_off(defer_cap, 0), read_pool_off(n_reads + 1, 0);
	std::vector<cs_mem_t> thread_mems((size_t)mem_cap * nthreads), pool(cap);
	const uint32_t kp1 = (uint32_t)opt->min_seed_len + 1;
	std::vector<cs_mem_t> r3_mems((size_t)n_bases / kp1 + n_reads + 2);
	std::vector<uint32_t> r3_n(n_reads + 1, 0);
	SeedArgs a;
	memset(&a, 0, sizeof a);
	a.bases = pb.data(); a.off = off; a.n_reads = n_reads; a.opt = *opt; a.packed = packed.data(); a.off_bias = 0; a.nmask = nmask.data();
	a.next_read = ctrl.data(); a.defer_bits = defer_bits.data(); a.defer_lx = defer_lx.data(); a.thread_lx = thread_lx.data();
	a.lit_q = lit_q.data(); a.n_lit = &n_lit; a.defer_cap = defer_cap; a.n_defer = &n_defer; a.n_defer_fast = &n_defer_fast;
	a.walk_order = walk_order.data(); a.n_walk = &n_walk; a.read_last_q = read_last_q.data();
	a.x_off = x_off.data(); a.x_n = x_n.data(); a.thread_mems = thread_mems.data(); a.mem_cap = mem_cap; a.spill = spill.data(); a.spill_cap = spill_cap;
	a.pool = pool.data(); a.pool_cap = cap; a.pool_used = &pool_used; a.read_pool_off = read_pool_off.data(); a.read_n_mems = read_n_mems.data();
	a.r3_mems = r3_mems.data(); a.r3_n_mems = r3_n.data();
	a.counters = counters.data(); a.req = req.data(); a.error = &error;
	const SeedArgs *pa = &a;
	if (mode == 0) {
		a.defer_q = defer_q.data();
		emul_warp([&]() { k_seed_fast(d, *pa); });
		n_defer_fast = n_defer;
		if (n_defer > defer_cap) return -102;
		uint32_t hist[64] = {0}, cursor[64] = {0};
		for (uint32_t qq = 0; qq < n_defer_fast; ++qq) if (defer_q[qq].y >> 31) ++hist[walk_cost_class(defer_q[qq], defer_bits[qq], (int)d.pt_k, (int)d.kt_depth)];
		k_walk_scan(hist, cursor, &n_walk);
		for (uint32_t qq = 0; qq < n_defer_fast; ++qq) if (defer_q[qq].y >> 31) walk_order[cursor[walk_cost_class(defer_q[qq], defer_bits[qq], (int)d.pt_k, (int)d.kt_depth)]++] = qq;
		emul_warp([&]() { k_seed_walk(d, *pa); });
		emul_warp([&]() { k_seed(d, *pa); });                       // call mode: the tasks listed in lit_q
		if (opt->max_mem_intv > 0) emul_warp([&]() { k_seed_r3_fast(d, *pa); });
	} else {
		a.defer_q = nullptr;
		emul_warp([&]() { k_seed(d, *pa); });                       // read mode: every read, literally
		if (opt->max_mem_intv > 0) emul_warp([&]() { k_seed_r3(d, *pa); });
	}
	if (error) return error;
	std::vector<std::vector<cs_mem_t>> per(n_reads);
	for (uint32_t r = 0; r < n_reads; ++r) {
		for (uint32_t m = 0; m < read_n_mems[r]; ++m) per[r].push_back(pool[read_pool_off[r] + m]);
		if (opt->max_mem_intv > 0) for (uint32_t m = 0; m < r3_n[r]; ++m) per[r].push_back(r3_mems[(size_t)(off[r] / kp1) + r + m]);
	}
	if (mode == 0)
		for (uint32_t q = 0; q < n_defer && q < defer_cap; ++q) {
			if (x_n[q] == 0xffffffffu) return -104;                  // a queued call nobody executed
			for (uint32_t m = 0; m < x_n[q]; ++m) per[defer_q[q].x].push_back(pool[x_off[q] + m]);
		}
	uint64_t n = 0;
	mem_off[0] = 0;
	for (uint32_t r = 0; r < n_reads; ++r) {
		std::stable_sort(per[r].begin(), per[r].end(), [](const cs_mem_t &x, const cs_mem_t &y) { return x.info < y.info; });
		if (n + per[r].size() > cap) return -100;
		for (const cs_mem_t &m : per[r]) out_mems[n++] = m;
		mem_off[r + 1] = (uint32_t)n;
	}
	if (stats) { for (int k = 0; k < 4; ++k) stats[k] = counters[k]; stats[4] = req[0]; stats[5] = req[1]; stats[6] = n_defer; stats[7] = n_lit; stats[12] = req[3]; stats[13] = req[2]; }
	return (int64_t)n;
}

}
