// Chaining and chain filtering on the device: mem_chain + mem_chain_flt of the reference
// (mapping/bwamem.c:359-497 == mapping/comp_seed.cpp:241-354), consuming the sorted mems and the resolved seed positions
// where k_collect_sort / k_sa_resolve left them, so that only chains cross the device-to-host link.
//
// The reference's result depends on ORDER in three places, and each is reproduced literally:
//   1. seeds are offered to the chains in emission order (bwamem.c:386-399), each to the chain kb_intervalp finds -- the
//      closest chain at or below its reference position in a B-tree keyed by the position of a chain's first seed
//      (KBTREE_INIT(chn, ...), cstl/kbtree.h).  Several chains may share a position (a repeat in the read against one
//      copy in the reference); which of them a lookup returns, and where a new equal key is inserted, depends on the shape
//      of the tree.  So the tree here IS that B-tree: t = 5 (KB_DEFAULT_SIZE 512 with a 40-byte mem_chain_t,
//      kbtree.h:54-60), lower-bound search inside a node (kbtree.h:117-131), insertion after the first key found
//      (:192-199), pre-emptive splits on the way down (:200-206, :174-190), in-order traversal (:336-359);
//   2. mem_chain_flt sorts the chains by weight with ks_introsort (cstl/ksort.h:176-226: median of three, partitions
//      down to 16 elements, then one insertion sort; comb sort when the depth budget runs out) -- not stable, so chains
//      of equal weight come out in the order that algorithm leaves them;
//   3. the pairwise overlap filter walks the kept chains in that order (bwamem.c:460-487).
// One read per thread (the work of a read is a chain of dependent steps), reads handed out by a work counter.
#include "cs_chain.cuh"

#define KB_T 5                     // minimum degree: a node holds at most 2t - 1 = 9 keys and 10 children
#define KB_MAXK (2 * KB_T - 1)
#define NODE_WORDS 24              // u32 per node: [0] n | is_internal << 31, [1..9] keys (chain ids), [10..19] children (node ids)
#define NONE 0xffffffffu

namespace {

struct Tree {
	uint32_t *nodes;    // the read's node region
	uint32_t n_nodes;   // allocated so far
	uint32_t cap;       // capacity of the region
	uint32_t root;
	bool overflow;
};

__device__ __forceinline__ uint32_t nd_n(const Tree &T, uint32_t x) { return T.nodes[x * NODE_WORDS] & 0x7fffffffu; }
__device__ __forceinline__ bool nd_internal(const Tree &T, uint32_t x) { return (T.nodes[x * NODE_WORDS] >> 31) != 0; }
__device__ __forceinline__ void nd_set(Tree &T, uint32_t x, uint32_t n, bool internal) { T.nodes[x * NODE_WORDS] = n | (internal ? 0x80000000u : 0u); }
__device__ __forceinline__ uint32_t &nd_key(Tree &T, uint32_t x, int i) { return T.nodes[x * NODE_WORDS + 1 + i]; }
__device__ __forceinline__ uint32_t &nd_ptr(Tree &T, uint32_t x, int i) { return T.nodes[x * NODE_WORDS + 1 + KB_MAXK + i]; }

__device__ __forceinline__ uint32_t nd_new(Tree &T, bool internal)
{
	if (T.n_nodes >= T.cap) { T.overflow = true; return 0; }
	const uint32_t z = T.n_nodes++;
	nd_set(T, z, 0, internal);
	return z;
}

} // namespace

struct ChainCtx { // what one thread needs to compare chains and seeds of its read
	const ChainArgs &a;
	uint64_t so;        // first seed of the read in the per-seed arrays
	__device__ __forceinline__ int64_t rb(uint32_t j) const { return (int64_t)a.rbeg[so + j]; }
	__device__ __forceinline__ int32_t qb(uint32_t j) const { return (int32_t)(a.s_qb_len[so + j] >> 16); }
	__device__ __forceinline__ int32_t ln(uint32_t j) const { return (int32_t)(a.s_qb_len[so + j] & 0xffffu); }
	__device__ __forceinline__ ChainTmp &ch(uint32_t c) const { return a.chains[so + c]; }
	__device__ __forceinline__ int64_t pos(uint32_t c) const { return rb(ch(c).first); }   // mem_chain_t.pos: position of the first seed
};

// chain_cmp(a, b) with b given by its position (bwamem.c:292)
__device__ __forceinline__ int cmp_pos(int64_t a, int64_t b) { return (int)(b < a) - (int)(a < b); }

// __kb_getp_aux (kbtree.h:117-131): index of the last key <= k in node x (-1 if none), *r = sign of (k - that key)... as the reference defines it
__device__ __forceinline__ int kb_getp_aux(Tree &T, const ChainCtx &C, uint32_t x, int64_t k, int *r)
{
	int begin = 0, end = (int)nd_n(T, x);
	const int n = end;
	if (n == 0) return -1;
	while (begin < end) {
		const int mid = (begin + end) >> 1;
		if (cmp_pos(C.pos(nd_key(T, x, mid)), k) < 0) begin = mid + 1;
		else end = mid;
	}
	if (begin == n) { *r = 1; return n - 1; }
	if ((*r = cmp_pos(k, C.pos(nd_key(T, x, begin)))) < 0) --begin;
	return begin;
}

// kb_intervalp (kbtree.h:151-169): the chain "lower" a seed at position k is offered to (NONE if there is none)
__device__ __forceinline__ uint32_t kb_lower(Tree &T, const ChainCtx &C, int64_t k)
{
	uint32_t x = T.root, lower = NONE;
	for (;;) {
		int r = 0;
		const int i = kb_getp_aux(T, C, x, k, &r);
		if (i >= 0 && r == 0) return nd_key(T, x, i);
		if (i >= 0) lower = nd_key(T, x, i);
		if (!nd_internal(T, x)) return lower;
		x = nd_ptr(T, x, i + 1);
	}
}

// __kb_split (kbtree.h:174-190): child y = ptr[i] of x is full; its upper half moves to a new node z = ptr[i + 1]
__device__ __forceinline__ void kb_split(Tree &T, uint32_t x, int i, uint32_t y)
{
	const bool yi = nd_internal(T, y);
	const uint32_t z = nd_new(T, yi);
	if (T.overflow) return;
	for (int k = 0; k < KB_T - 1; ++k) nd_key(T, z, k) = nd_key(T, y, KB_T + k);
	if (yi) for (int k = 0; k < KB_T; ++k) nd_ptr(T, z, k) = nd_ptr(T, y, KB_T + k);
	nd_set(T, z, KB_T - 1, yi);
	nd_set(T, y, KB_T - 1, yi);
	const int xn = (int)nd_n(T, x);
	for (int k = xn; k > i; --k) nd_ptr(T, x, k + 1) = nd_ptr(T, x, k);
	nd_ptr(T, x, i + 1) = z;
	for (int k = xn - 1; k >= i; --k) nd_key(T, x, k + 1) = nd_key(T, x, k);
	nd_key(T, x, i) = nd_key(T, y, KB_T - 1);
	nd_set(T, x, (uint32_t)xn + 1, true);
}

// kb_putp (kbtree.h:208-221) + __kb_putp_aux (:191-207): insert chain c (position k)
__device__ __forceinline__ void kb_put(Tree &T, const ChainCtx &C, uint32_t c, int64_t k)
{
	uint32_t r = T.root;
	if (nd_n(T, r) == KB_MAXK) {
		const uint32_t s = nd_new(T, true);
		if (T.overflow) return;
		T.root = s;
		nd_ptr(T, s, 0) = r;
		kb_split(T, s, 0, r);
		if (T.overflow) return;
		r = s;
	}
	uint32_t x = r;
	for (;;) {
		int rr;
		if (!nd_internal(T, x)) {
			const int i = kb_getp_aux(T, C, x, k, &rr), n = (int)nd_n(T, x);
			for (int q = n - 1; q > i; --q) nd_key(T, x, q + 1) = nd_key(T, x, q);
			nd_key(T, x, i + 1) = c;
			nd_set(T, x, (uint32_t)n + 1, false);
			return;
		}
		int i = kb_getp_aux(T, C, x, k, &rr) + 1;
		if (nd_n(T, nd_ptr(T, x, i)) == KB_MAXK) {
			kb_split(T, x, i, nd_ptr(T, x, i));
			if (T.overflow) return;
			if (cmp_pos(k, C.pos(nd_key(T, x, i))) > 0) ++i;
		}
		x = nd_ptr(T, x, i);
	}
}

// bns_pos2rid (bntseq.c:354-368) / bns_intv2rid (:370-378)
// hint: the sequence the caller's previous position fell into (the seeds of a read mostly lie in one, and both ends of a seed
// almost always do): tried first, instead of the binary search (20 % of the kernel's instructions, profiles/r02_ncu_k_chain_build_*)
__device__ __forceinline__ int pos2rid(const ChainArgs &a, int64_t pos_f, int hint = -1)
{
	if (pos_f >= a.l_pac) return -1;
	if (hint >= 0 && hint < a.n_seqs && pos_f >= a.c_off[hint] && (hint == a.n_seqs - 1 || pos_f < a.c_off[hint + 1])) return hint;
	int left = 0, mid = 0, right = a.n_seqs;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= a.c_off[mid]) {
			if (mid == a.n_seqs - 1) break;
			if (pos_f < a.c_off[mid + 1]) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}
__device__ __forceinline__ int64_t depos(const ChainArgs &a, int64_t pos) { return pos >= a.l_pac ? (a.l_pac << 1) - 1 - pos : pos; }
__device__ __forceinline__ int intv2rid(const ChainArgs &a, int64_t rb, int64_t re, int hint = -1)
{
	if (rb < a.l_pac && re > a.l_pac) return -2;
	const int rid_b = pos2rid(a, depos(a, rb), hint);
	const int rid_e = rb < re ? pos2rid(a, depos(a, re - 1), rid_b) : rid_b;
	return rid_b == rid_e ? rid_b : -1;
}

// test_and_merge (bwamem.c:296-320): 1 if seed j went into chain c (or is contained in it)
__device__ __forceinline__ int test_and_merge(const ChainArgs &a, const ChainCtx &C, uint32_t c, uint32_t j, int seed_rid)
{
	ChainTmp &ch = C.ch(c);
	const uint32_t f = ch.first, l = ch.last;
	const int64_t qend = (int64_t)C.qb(l) + C.ln(l), rend = C.rb(l) + C.ln(l);
	const int64_t p_rb = C.rb(j); const int32_t p_qb = C.qb(j), p_ln = C.ln(j);
	if (seed_rid != ch.rid) return 0;
	if (p_qb >= C.qb(f) && (int64_t)p_qb + p_ln <= qend && p_rb >= C.rb(f) && p_rb + p_ln <= rend) return 1;
	if ((C.rb(l) < a.l_pac || C.rb(f) < a.l_pac) && p_rb >= a.l_pac) return 0;
	const int64_t x = (int64_t)p_qb - C.qb(l), y = p_rb - C.rb(l);
	if (y >= 0 && x - y <= a.copt.w && y - x <= a.copt.w && x - C.ln(l) < a.copt.max_chain_gap && y - C.ln(l) < a.copt.max_chain_gap) {
		a.s_next[C.so + l] = j;
		ch.last = j; ++ch.n;
		return 1;
	}
	return 0;
}

// mem_chain_weight (bwamem.c:322-343)
__device__ __forceinline__ uint32_t chain_weight(const ChainArgs &a, const ChainCtx &C, uint32_t c)
{
	const ChainTmp &ch = C.ch(c);
	int64_t end = 0; int w = 0, tmp;
	uint32_t j = ch.first;
	for (uint32_t k = 0; k < ch.n; ++k, j = a.s_next[C.so + j]) {
		const int64_t qb = C.qb(j), ln = C.ln(j);
		if (qb >= end) w += (int)ln;
		else if (qb + ln > end) w += (int)(qb + ln - end);
		end = end > qb + ln ? end : qb + ln;
	}
	tmp = w; w = 0; end = 0; j = ch.first;
	for (uint32_t k = 0; k < ch.n; ++k, j = a.s_next[C.so + j]) {
		const int64_t rb = C.rb(j), ln = C.ln(j);
		if (rb >= end) w += (int)ln;
		else if (rb + ln > end) w += (int)(rb + ln - end);
		end = end > rb + ln ? end : rb + ln;
	}
	w = w < tmp ? w : tmp;
	return (uint32_t)(w < (1 << 30) ? w : (1 << 30) - 1);
}

// ---- ks_introsort(mem_flt) on an array of chain ids; flt_lt(a, b) = a.w > b.w (bwamem.c:432-433, ksort.h:146-226) ----
#define LT(p, q) (C.ch(p).w > C.ch(q).w)

__device__ __forceinline__ void ks_insertsort(const ChainCtx &C, uint32_t *s, uint32_t *t)
{
	for (uint32_t *i = s + 1; i < t; ++i)
		for (uint32_t *j = i; j > s && LT(*j, *(j - 1)); --j) { const uint32_t tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

__device__ void ks_combsort(const ChainCtx &C, size_t n, uint32_t *a)
{
	const double shrink_factor = 1.2473309501039786540366528676643;
	int do_swap;
	size_t gap = n;
	do {
		if (gap > 2) {
			gap = (size_t)(gap / shrink_factor);
			if (gap == 9 || gap == 10) gap = 11;
		}
		do_swap = 0;
		for (uint32_t *i = a; i < a + n - gap; ++i) {
			uint32_t *j = i + gap;
			if (LT(*j, *i)) { const uint32_t tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
		}
	} while (do_swap || gap > 2);
	if (gap != 1) ks_insertsort(C, a, a + n);
}

__device__ void ks_introsort(const ChainCtx &C, size_t n, uint32_t *a)
{
	int d;
	uint32_t *st_l[70], *st_r[70]; int st_d[70]; int top = 0;      // (sizeof(size_t) * d) + 2 entries in the reference; d <= 32 here
	uint32_t rp, tmp;
	uint32_t *s, *t, *i, *j, *k;
	if (n < 1) return;
	else if (n == 2) {
		if (LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; }
		return;
	}
	for (d = 2; (1ul << d) < n; ++d);
	s = a; t = a + (n - 1); d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) {
				ks_combsort(C, (size_t)(t - s) + 1, s);
				t = s;
				continue;
			}
			i = s; j = t; k = i + ((j - i) >> 1) + 1;
			if (LT(*k, *i)) {
				if (LT(*k, *j)) k = j;
			} else k = LT(*j, *i) ? i : j;
			rp = *k;
			if (k != t) { tmp = *k; *k = *t; *t = tmp; }
			for (;;) {
				do ++i; while (LT(*i, rp));
				do --j; while (i <= j && LT(rp, *j));
				if (j <= i) break;
				tmp = *i; *i = *j; *j = tmp;
			}
			tmp = *i; *i = *t; *t = tmp;
			if (i - s > t - i) {
				if (i - s > 16) { st_l[top] = s; st_r[top] = i - 1; st_d[top] = d; ++top; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { st_l[top] = i + 1; st_r[top] = t; st_d[top] = d; ++top; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == 0) { ks_insertsort(C, a, a + n); return; }
			--top; s = st_l[top]; t = st_r[top]; d = st_d[top];
		}
	}
}
#undef LT

// ---------------------------------------------------------------------------------------------
// Pass 1: chains of every read, filtered; leaves per read the kept chains (ids, in output order) in order[] and their
// count / seed total for the offsets scan.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chain_build(ChainArgs a)
{
	for (;;) {
		const unsigned long long rr = atomicAdd(a.work, 1ull);
		if (rr >= a.n_reads) break;
		const uint32_t r = (uint32_t)rr;
		const uint64_t so = a.seed_off[r];
		const uint32_t S = a.seed_off[r + 1] - a.seed_off[r];
		const int len = (int)(a.off[r + 1] - a.off[r]);
		const cs_mem_t *mem = a.mems + a.mem_off[r];
		const uint32_t n_mem = a.mem_off[r + 1] - a.mem_off[r];
		const ChainCtx C{a, so};
		uint32_t n_out = 0, seeds_out = 0, l_rep = 0;
		a.n_chain[r] = 0; a.n_cseed[r] = 0; a.l_rep[r] = 0;
		if (len < a.opt.min_seed_len || S == 0) continue;           // bwamem.c:368: no match for a query shorter than the seed length
		if ((uint64_t)a.seed_off[r + 1] > a.seed_cap || (uint64_t)a.mem_off[r + 1] > a.mems_cap) continue;   // the batch overflowed its buffers: reported by the collect pass
		if ((uint64_t)a.node_off[r + 1] > a.node_cap) { atomicMin(a.error, CS_E_OVERFLOW); continue; }
		{ // fraction of the read covered by repetitive seeds (bwamem.c:377-385)
			int b = 0, e = 0, lr = 0;
			for (uint32_t i = 0; i < n_mem; ++i) {
				const uint64_t info = mem[i].info;
				const int sb = (int)(info >> 32), se = (int)(uint32_t)info;
				if (mem[i].x[2] <= (uint64_t)a.opt.max_occ) continue;
				if (sb > e) { lr += e - b; b = sb; e = se; }
				else e = e > se ? e : se;
			}
			lr += e - b;
			l_rep = (uint32_t)lr;
		}
		Tree T;
		T.nodes = a.nodes + (size_t)a.node_off[r] * NODE_WORDS; T.cap = a.node_off[r + 1] - a.node_off[r];
		T.n_nodes = 0; T.overflow = false;
		T.root = nd_new(T, false);
		uint32_t n_ch = 0, j = 0;
		int last_rid = -1;                                          // (hint for the sequence look-up of the next seed)
		for (uint32_t i = 0; i < n_mem; ++i) { // seeds in emission order (bwamem.c:386-399)
			const uint64_t info = mem[i].info, x2 = mem[i].x[2];
			const int32_t qbeg = (int32_t)(info >> 32), slen = (int32_t)(uint32_t)info - qbeg;
			const uint32_t cnt = x2 < (uint64_t)a.opt.max_occ ? (uint32_t)x2 : (uint32_t)a.opt.max_occ;
			for (uint32_t k = 0; k < cnt; ++k, ++j) {
				a.s_qb_len[so + j] = ((uint32_t)qbeg << 16) | (uint32_t)slen;
				a.s_next[so + j] = NONE;
				const int64_t rb = C.rb(j);
				const int rid = intv2rid(a, rb, rb + slen, last_rid);
				if (rid >= 0) last_rid = rid;
				if (rid < 0) continue;                                  // bridging two sequences or the strand boundary (bwamem.c:403)
				bool merged = false;
				if (n_ch) {
					const uint32_t lower = kb_lower(T, C, rb);
					if (lower != NONE && test_and_merge(a, C, lower, j, rid)) merged = true;
				}
				if (!merged) {
					ChainTmp &c = C.ch(n_ch);
					c.first = c.last = j; c.n = 1; c.rid = rid; c.w = 0; c.first_sh = -1; c.kept = 0;
					kb_put(T, C, n_ch, rb);
					++n_ch;
				}
			}
		}
		if (T.overflow) { atomicMin(a.error, CS_E_READ_OVERFLOW); continue; }   // (cannot happen: the region is sized for the worst tree)
		uint32_t *ord = a.order + so;
		{ // __kb_traverse (kbtree.h:336-359): chains in key order
			uint32_t st_x[18]; int st_i[18]; int sp = 0; uint32_t n = 0;
			st_x[0] = T.root; st_i[0] = 0;
			for (;;) {
				while (st_x[sp] != NONE && st_i[sp] <= (int)nd_n(T, st_x[sp])) {
					const uint32_t child = nd_internal(T, st_x[sp]) ? nd_ptr(T, st_x[sp], st_i[sp]) : NONE;
					++sp; st_x[sp] = child; st_i[sp] = 0;
				}
				--sp;
				if (sp < 0) break;
				if (st_x[sp] != NONE && st_i[sp] < (int)nd_n(T, st_x[sp])) ord[n++] = nd_key(T, st_x[sp], st_i[sp]);
				++st_i[sp];
			}
		}
		// mem_chain_flt (bwamem.c:435-497)
		uint32_t n_chn = 0;
		for (uint32_t i = 0; i < n_ch; ++i) {
			const uint32_t c = ord[i];
			const uint32_t w = chain_weight(a, C, c);
			C.ch(c).w = w;
			if ((int)w >= a.copt.min_chain_weight) ord[n_chn++] = c;
		}
		if (n_chn) {
			ks_introsort(C, n_chn, ord);
			uint32_t *kl = a.klist + so; uint32_t nk = 0;
#define CB(c) (C.qb(C.ch(c).first))
#define CE(c) (C.qb(C.ch(c).last) + C.ln(C.ch(c).last))
			C.ch(ord[0]).kept = 3; kl[nk++] = 0;
			for (uint32_t i = 1; i < n_chn; ++i) {
				int large_ovlp = 0;
				uint32_t k;
				const uint32_t ci = ord[i];
				const int bi = CB(ci), ei = CE(ci);
				const bool alt_i = a.c_alt && a.c_alt[C.ch(ci).rid];
				for (k = 0; k < nk; ++k) {
					const uint32_t jx = kl[k], cj = ord[jx];
					const int bj = CB(cj), ej = CE(cj);
					const int b_max = bj > bi ? bj : bi, e_min = ej < ei ? ej : ei;
					const bool alt_j = a.c_alt && a.c_alt[C.ch(cj).rid];
					if (e_min > b_max && (!alt_j || alt_i)) {
						const int li = ei - bi, lj = ej - bj, min_l = li < lj ? li : lj;
						if ((float)(e_min - b_max) >= (float)min_l * a.copt.mask_level && min_l < a.copt.max_chain_gap) {
							large_ovlp = 1;
							if (C.ch(cj).first_sh < 0) C.ch(cj).first_sh = (int32_t)i;
							if ((float)(int)C.ch(ci).w < (float)(int)C.ch(cj).w * a.copt.drop_ratio && (int)C.ch(cj).w - (int)C.ch(ci).w >= (a.opt.min_seed_len << 1)) break;
						}
					}
				}
				if (k == nk) { kl[nk++] = i; C.ch(ci).kept = large_ovlp ? 2 : 3; }
			}
#undef CB
#undef CE
			for (uint32_t i = 0; i < nk; ++i) {
				const ChainTmp &c = C.ch(ord[kl[i]]);
				if (c.first_sh >= 0) C.ch(ord[c.first_sh]).kept = 1;
			}
			uint32_t i = 0; int k2 = 0;
			for (; i < n_chn; ++i) { // don't extend more than max_chain_extend .kept=1/2 chains
				const uint32_t kp = C.ch(ord[i]).kept;
				if (kp == 0 || kp == 3) continue;
				if (++k2 >= a.copt.max_chain_extend) break;
			}
			for (; i < n_chn; ++i) if (C.ch(ord[i]).kept < 3) C.ch(ord[i]).kept = 0;
			for (i = 0; i < n_chn; ++i) {
				const uint32_t c = ord[i];
				if (C.ch(c).kept == 0) continue;
				ord[n_out++] = c; seeds_out += C.ch(c).n;
			}
		}
		a.n_chain[r] = n_out; a.n_cseed[r] = seeds_out; a.l_rep[r] = l_rep;
	}
}

// Pass 2: chain records and their seeds at the offsets the scans gave.  Eight lanes per read.
__global__ void k_chain_emit(ChainArgs a, const uint32_t *chain_off, const uint32_t *cseed_off, uint64_t chain_cap, uint64_t cseed_cap,
                             cs_chain_t *out, uint32_t *s_lo, uint8_t *s_hi, uint16_t *s_qbeg, uint16_t *s_len)
{
	const uint32_t sub = threadIdx.x & 7;
	const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
	for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < a.n_reads; r += ngroups) {
		const uint32_t n = chain_off[r + 1] - chain_off[r];
		if (n == 0) continue;
		if ((uint64_t)chain_off[r + 1] > chain_cap || (uint64_t)cseed_off[r + 1] > cseed_cap) { if (sub == 0) atomicMin(a.error, CS_E_OVERFLOW); continue; }
		const uint64_t so = a.seed_off[r];
		const ChainCtx C{a, so};
		const uint32_t *ord = a.order + so;
		uint64_t o = cseed_off[r];
		for (uint32_t i = 0; i < n; ++i) {
			const ChainTmp &c = C.ch(ord[i]);
			if (sub == 0) {
				cs_chain_t rec;
				rec.rid = c.rid;
				rec.w_kept = (c.w & 0x1fffffffu) | (c.kept << 29) | ((a.c_alt && a.c_alt[c.rid]) ? 0x80000000u : 0u);
				rec.n = c.n; rec.l_rep = a.l_rep[r];
				out[chain_off[r] + i] = rec;
			}
			uint32_t j = c.first;
			for (uint32_t k = 0; k < c.n; ++k, j = a.s_next[so + j]) { // (a linked list: walked by every lane, written by one of eight)
				if ((k & 7) != sub) continue;
				const uint64_t rb = a.rbeg[so + j];
				s_lo[o + k] = (uint32_t)rb; s_hi[o + k] = (uint8_t)(rb >> 32);
				s_qbeg[o + k] = (uint16_t)C.qb(j); s_len[o + k] = (uint16_t)C.ln(j);
			}
			o += c.n;
		}
	}
}

// node capacity of each read's B-tree region: nine keys fit in the root; beyond that every node but the root holds at least
// t - 1 = 4 of the K <= S keys (splits leave exactly 4 on each side, kbtree.h:174-190), so there are at most (K - 1) / 4 + 1 nodes.
// The sum over a batch is at most 2 * n_reads + n_seeds / 4: what cs_ctx_set_chaining allocates (node_cap).
__global__ void k_chain_node_counts(const uint32_t *read_n_seeds, uint32_t n_reads, uint32_t *out)
{
	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
		const uint32_t S = read_n_seeds[r];
		out[r] = S <= KB_MAXK ? 1u : S / 4 + 2;
	}
}
