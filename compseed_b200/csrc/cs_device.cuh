// Device-side FM-index primitives for sm_100a.
//
// HBM layout of the index (built once at upload by k_relayout from the reference's 64-byte
// buckets, FM_index/bwt.h:74-80, index_main.c:152-174):
//
//   bucket b (32 bytes, one L2/DRAM sector, 64 BWT rows [64b, 64b+64)):
//     u32 w0..w3   bases LSB-first: row 64b+i is at bits 2*(i&15) of w[i>>4]
//     u32 p1, p2, p3   low 32 bits of the PREFIX sums #A, #A+#C, #A+#C+#G over rows [0, 64b)
//     u32 hi           bits 0-7 / 8-15 / 16-23: bits 32-39 of p1 / p2 / p3
//   The last prefix sum is implied: p4 = 64b (the '$' row is not stored, as in the reference).
//   Prefix sums let one extend pick "bases <= c" and "bases < c" with two selects instead of
//   assembling four 64-bit counts.  One occ lookup == one 256-bit load (LDG.E.256) == one 32-byte sector.
#pragma once
#include <cstdint>
#ifndef CS_EMUL   // tests/emul/seed_emul.cpp compiles this file as plain C++ (test infrastructure; never part of the library)
#include <cuda_runtime.h>
#endif

struct DevIndex {
	const uint4 *buckets;   // 2 x uint4 per bucket
	uint64_t n_buckets;
	uint64_t primary, seq_len;
	uint64_t L2[5];
	const uint64_t *sa;     // sa[0] == (uint64_t)-1
	uint64_t n_sa;
	uint32_t sa_mask;       // sa_intv - 1
	uint32_t sa_shift;      // log2(sa_intv)
	// Top-of-search table (the SST's role, SURVEY section 7 hard part 2): the bi-interval of EVERY
	// string of 1..kt_depth bases, 16 bytes each, depth d at entry offset (4^d - 4)/3.  A pure memo of
	// bwt_extend: entry(S + b) == bwt_extend(entry(S), forward, b).  Key: base j of the string at bits 2j.
	const uint4 *kt;
	uint32_t kt_depth;      // 0 = no table
	// Occurrence filter: a 2-bit saturating count (0,1,2,>=3) of every pt_k-mer of the indexed text
	// (both strands), key as in read_key.  Lets the backward phase drop, with one gather, every forward
	// match that cannot grow to min_seed_len bases (see "Occurrence filter" at k_seed).  0 = no filter.
	const uint32_t *pt;
	uint32_t pt_k;
	// Unique-match fast path (k_seed, ST_TXT_*): the indexed text T = fwd + revcomp(fwd), 2 bits per
	// base, 32 bases per u64 with base j of a word at bits 2j (like the packed reads), and the inverse
	// suffix array sampled every 2^isa_shift text positions (isa[p >> isa_shift] = row of suffix p).
	// Needs the dense SA (sa_mask == 0).  text == NULL: fast path off.
	const uint64_t *text;
	const uint64_t *isa;
	uint32_t isa_shift;
	// Repeat lengths (k_rep_build): rep[p] = the largest d such that T[p, p+d) occurs at least twice in T (as a prefix of
	// two suffixes), capped at 255 = "255 or more".  With it a bwt_smem1a call that starts at a KNOWN text position (the
	// second-pass call in the middle of a one-occurrence SMEM, bwamem.c:238-249) needs neither the FM-index nor the
	// occurrence filter: its forward match is rep[p] bases long, and the K-mer windows inside the SMEM occur twice iff
	// their rep is >= K.  One byte per text position.  NULL: not built (needs the dense SA and the text).
	const uint8_t *rep;
};


// Random single-word gathers from the big tables (filter, SA, inverse SA, text, top-of-search table).  ncu
// shows L2 filling about four sectors per such request by default (lts__t_sectors_srcunit_tex_op_read ~ 3.7 x
// lts__t_requests); nothing else of the 128-byte line is ever used, so the prefetch size is capped at 64 bytes,
// the smallest PTX offers.  -DCS_L2_DEFAULT restores plain __ldg.
#if defined(CS_L2_DEFAULT) || defined(CS_EMUL)
__device__ __forceinline__ uint32_t gather_u32(const uint32_t *p) { return __ldg(p); }
__device__ __forceinline__ uint64_t gather_u64(const uint64_t *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t gather_u8(const uint8_t *p) { return __ldg(p); }
__device__ __forceinline__ uint4 gather_u128(const uint4 *p) { return __ldg(p); }
#else
__device__ __forceinline__ uint32_t gather_u32(const uint32_t *p)
{ uint32_t v; asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint64_t gather_u64(const uint64_t *p)
{ uint64_t v; asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t gather_u8(const uint8_t *p)
{ uint32_t v; asm volatile("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint4 gather_u128(const uint4 *p)
{ uint4 v; asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
#endif

struct Bucket { uint32_t w0, w1, w2, w3; uint32_t p1, p2, p3, hi; };

// L2[c] for a run-time c without indexing the kernel-parameter struct dynamically (which would
// force a local-memory copy of it)
__device__ __forceinline__ uint64_t l2_at(const DevIndex &I, int c)
{
	uint64_t lo = (c & 1) ? I.L2[1] : I.L2[0];
	uint64_t hi = (c & 1) ? I.L2[3] : I.L2[2];
	return c == 4 ? I.L2[4] : ((c & 2) ? hi : lo);
}

__device__ __forceinline__ Bucket load_bucket(const DevIndex &I, uint64_t b)
{
	Bucket r;
	const uint4 *p = I.buckets + 2 * b;
#ifdef CS_EMUL
	r.w0 = p[0].x; r.w1 = p[0].y; r.w2 = p[0].z; r.w3 = p[0].w; r.p1 = p[1].x; r.p2 = p[1].y; r.p3 = p[1].z; r.hi = p[1].w;
#else
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3), "=r"(r.p1), "=r"(r.p2), "=r"(r.p3), "=r"(r.hi) : "l"(p));
#endif
	return r;
}

// Low-bit-plane masks of the first n bases of a bucket (n = 0..64), one 32-bit mask per 16-base word.
// 1 KB, always L1-resident; fetched together with the bucket so the masks cost no dependent ALU work.
__device__ const uint4 g_occ_mask[65] = {
	{0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00000005u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00000015u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00000055u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00000155u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00000555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00001555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00005555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00015555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00055555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00155555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x00555555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x01555555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x05555555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x15555555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000000u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000001u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000005u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000015u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000055u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000155u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00000555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00001555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00005555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00015555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00055555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00155555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x00555555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x01555555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x05555555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x15555555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000000u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000001u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000005u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000015u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000055u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000155u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00000555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00001555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00005555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00015555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00055555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00155555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x00555555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x01555555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x05555555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x15555555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000000u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000001u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000005u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000015u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000055u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000155u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00000555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00001555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00005555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00015555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00055555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00155555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x00555555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x01555555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x05555555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x15555555u},
	{0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u}
};

// Base counts among the first n bases of a bucket from three popcount pairs:
//   T = #T, H = #G + #T (high bit set), E = #C + #T (low bit set)   =>   #A = n - H - E + T.
// Words are merged pairwise (word 0 on the even bits, word 1 shifted onto the odd bits) to halve the popcounts.
__device__ __forceinline__ void bucket_the(const Bucket &B, uint4 m, uint32_t &T, uint32_t &H, uint32_t &E)
{
	uint32_t m1s = m.y << 1, m3s = m.w << 1;
	uint32_t E01 = (B.w0 & m.x) | ((B.w1 << 1) & m1s), H01 = ((B.w0 >> 1) & m.x) | (B.w1 & m1s);
	uint32_t E23 = (B.w2 & m.z) | ((B.w3 << 1) & m3s), H23 = ((B.w2 >> 1) & m.z) | (B.w3 & m3s);
	E = __popc(E01) + __popc(E23);
	H = __popc(H01) + __popc(H23);
	T = __popc(E01 & H01) + __popc(E23 & H23);
}

// checkpoint prefix sums: P0 = 0, P1 = #A, P2 = #A+#C, P3 = #A+#C+#G, P4 = 64b (rows before the bucket)
__device__ __forceinline__ uint64_t bucket_p(const Bucket &B, int j) // j = 1..3
{
	uint32_t lo = j == 1 ? B.p1 : j == 2 ? B.p2 : B.p3;
	return (uint64_t)lo | ((uint64_t)((B.hi >> (8 * (j - 1))) & 0xff) << 32);
}

// the base stored at position r (0..63) of a bucket
__device__ __forceinline__ uint32_t bucket_base(const Bucket &B, uint32_t r)
{
	uint32_t w = (r & 32) ? ((r & 16) ? B.w3 : B.w2) : ((r & 16) ? B.w1 : B.w0);
	return (w >> (2 * (r & 15))) & 3;
}

// cnt[c] = number of base c in rows [0 .. 64b + r] of the stored BWT (r in 0..63, inclusive)
__device__ __forceinline__ void bucket_occ4(const Bucket &B, uint64_t b, uint32_t r, uint64_t cnt[4])
{
	uint32_t n = r + 1, T, H, E;
	bucket_the(B, __ldg(&g_occ_mask[n]), T, H, E);
	uint64_t P1 = bucket_p(B, 1), P2 = bucket_p(B, 2), P3 = bucket_p(B, 3);
	cnt[0] = P1 + (n - H - E + T);
	cnt[1] = (P2 - P1) + (E - T);
	cnt[2] = (P3 - P2) + (H - T);
	cnt[3] = ((b << 6) - P3) + T;
}

// bwt_occ4 (FM_index/bwt.c:169-186) for k != -1
__device__ __forceinline__ void dev_occ4(const DevIndex &I, uint64_t k, uint64_t cnt[4])
{
	k -= (k >= I.primary);
	uint64_t b = k >> 6;
	Bucket B = load_bucket(I, b);
	bucket_occ4(B, b, (uint32_t)k & 63, cnt);
}

// bwt_extend (FM_index/bwt.c:262-275), returning only child c.  in/out: (x0, x1, x2).
// Written around cumulative counts: with U(pos) = #bases <= c and W(pos) = #bases < c in rows [0..pos],
//   occ(pos, c) = U - W,   ok[c].x[2] = occ(l,c) - occ(k,c),
//   sum of ok[j].x[2] for j > c (what bwt.c:271-274 accumulates) = (l' - k') - (U(l) - U(k)),
// so only two 64-bit values per position are assembled, selected by c from the prefix-sum checkpoint
// and from (n, T, H, E) of the bucket.  `two` = k and l needed two different sectors.
__device__ __forceinline__ void dev_extend(const DevIndex &I, uint64_t x0, uint64_t x1, uint64_t x2, int c, int is_back,
                                           uint64_t &o0, uint64_t &o1, uint64_t &o2, uint32_t &two)
{
	const uint64_t a = is_back ? x0 : x1;      // x[!is_back]
	const uint64_t o = is_back ? x1 : x0;      // x[is_back]
	const uint64_t k = a - 1, l = k + x2;
	const uint64_t kk = k - (k >= I.primary), ll = l - (l >= I.primary);
	const uint64_t bk = kk >> 6, bl = ll >> 6;
	const uint32_t nk = ((uint32_t)kk & 63) + 1, nl = ((uint32_t)ll & 63) + 1;
	const bool same = (bl == bk);
	const uint4 mk = __ldg(&g_occ_mask[nk]), ml = __ldg(&g_occ_mask[nl]);
	Bucket Bk = load_bucket(I, bk), Bl;
	if (!same) Bl = load_bucket(I, bl); else Bl = Bk;
	two = !same;
	const bool c1 = c & 1, c2 = c & 2;
	uint64_t Uk, Wk, Ul, Wl;
	{
		uint32_t T, H, E;
		bucket_the(Bk, mk, T, H, E);
		const uint32_t v1 = nk - H - E + T, v2 = nk - H, v3 = nk - T;           // #A, #A+#C, #A+#C+#G among the first nk
		const uint32_t ui = c2 ? (c1 ? nk : v3) : (c1 ? v2 : v1), wi = c2 ? (c1 ? v3 : v2) : (c1 ? v1 : 0u);
		const uint32_t ulo = c2 ? (c1 ? (uint32_t)(bk << 6) : Bk.p3) : (c1 ? Bk.p2 : Bk.p1);
		const uint32_t wlo = c2 ? (c1 ? Bk.p3 : Bk.p2) : (c1 ? Bk.p1 : 0u);
		const uint32_t uhi = (c1 && c2) ? (uint32_t)(bk >> 26) : ((Bk.hi >> (8 * c)) & 0xff);
		const uint32_t whi = c ? ((Bk.hi >> (8 * (c - 1))) & 0xff) : 0u;
		Uk = (((uint64_t)uhi << 32) | ulo) + ui;
		Wk = (((uint64_t)whi << 32) | wlo) + wi;
	}
	{
		uint32_t T, H, E;
		bucket_the(Bl, ml, T, H, E);
		const uint32_t v1 = nl - H - E + T, v2 = nl - H, v3 = nl - T;
		const uint32_t ui = c2 ? (c1 ? nl : v3) : (c1 ? v2 : v1), wi = c2 ? (c1 ? v3 : v2) : (c1 ? v1 : 0u);
		const uint32_t ulo = c2 ? (c1 ? (uint32_t)(bl << 6) : Bl.p3) : (c1 ? Bl.p2 : Bl.p1);
		const uint32_t wlo = c2 ? (c1 ? Bl.p3 : Bl.p2) : (c1 ? Bl.p1 : 0u);
		const uint32_t uhi = (c1 && c2) ? (uint32_t)(bl >> 26) : ((Bl.hi >> (8 * c)) & 0xff);
		const uint32_t whi = c ? ((Bl.hi >> (8 * (c - 1))) & 0xff) : 0u;
		Ul = (((uint64_t)uhi << 32) | ulo) + ui;
		Wl = (((uint64_t)whi << 32) | wlo) + wi;
	}
	const uint64_t tk = Uk - Wk, tl = Ul - Wl;
	const uint64_t gsum = (ll - kk) - (Ul - Uk);
	const uint64_t base = o + ((a <= I.primary) && (a + x2 - 1 >= I.primary));
	const uint64_t na = l2_at(I, c) + 1 + tk, no = base + gsum;
	o0 = is_back ? na : no;
	o1 = is_back ? no : na;
	o2 = tl - tk;
}

// all four children (for the cs_extend probe): ok[c*3 + j]
__device__ __forceinline__ void dev_extend4(const DevIndex &I, const uint64_t ik[3], int is_back, uint64_t *ok)
{
	int a = !is_back, b = is_back;
	uint64_t tk[4], tl[4];
	uint64_t k = ik[a] - 1, l = k + ik[2];
	if (k == (uint64_t)-1) { tk[0] = tk[1] = tk[2] = tk[3] = 0; } else dev_occ4(I, k, tk);
	if (l == (uint64_t)-1) { tl[0] = tl[1] = tl[2] = tl[3] = 0; } else dev_occ4(I, l, tl);
	for (int i = 0; i < 4; ++i) { ok[i * 3 + a] = I.L2[i] + 1 + tk[i]; ok[i * 3 + 2] = tl[i] - tk[i]; }
	ok[3 * 3 + b] = ik[b] + (ik[a] <= I.primary && ik[a] + ik[2] - 1 >= I.primary);
	ok[2 * 3 + b] = ok[3 * 3 + b] + ok[3 * 3 + 2];
	ok[1 * 3 + b] = ok[2 * 3 + b] + ok[2 * 3 + 2];
	ok[0 * 3 + b] = ok[1 * 3 + b] + ok[1 * 3 + 2];
}

// one LF step: bwt_invPsi (FM_index/bwt.c:53-59) with bwt_occ (bwt.c:107-129) of the same bucket
__device__ __forceinline__ uint64_t dev_lf(const DevIndex &I, uint64_t k)
{
	if (k == I.primary) return 0;
	uint64_t x = k - (k > I.primary);          // position of row k's char in the stored BWT
	uint64_t b = x >> 6;
	uint32_t r = (uint32_t)x & 63;
	Bucket B = load_bucket(I, b);
	uint32_t c = bucket_base(B, r);
	// occ(k, c) counts rows [0..k] inclusive; k - (k >= primary) == x for k != primary
	uint64_t cnt[4];
	bucket_occ4(B, b, r, cnt);
	return l2_at(I, (int)c) + (c == 0 ? cnt[0] : c == 1 ? cnt[1] : c == 2 ? cnt[2] : cnt[3]);
}

// bwt_sa (FM_index/bwt.c:86-96)
__device__ __forceinline__ uint64_t dev_sa(const DevIndex &I, uint64_t k, uint32_t &steps)
{
	uint64_t sa = 0;
	while (k & I.sa_mask) { ++sa; k = dev_lf(I, k); }
	steps = (uint32_t)sa;
	return sa + gather_u64(I.sa + (k >> I.sa_shift));
}

// ---- top-of-search table and 2-bit packed reads ----
__device__ __forceinline__ uint64_t kt_offset(uint32_t d) { return ((1ull << (2 * d)) - 4) / 3; }

__device__ __forceinline__ void kt_lookup(const DevIndex &I, uint32_t d, uint64_t key, uint64_t &x0, uint64_t &x1, uint64_t &x2)
{
	uint4 v = gather_u128(I.kt + kt_offset(d) + key);
	x0 = (uint64_t)v.x | ((uint64_t)(v.w & 31) << 32);
	x1 = (uint64_t)v.y | ((uint64_t)((v.w >> 5) & 31) << 32);
	x2 = (uint64_t)v.z | ((uint64_t)((v.w >> 10) & 31) << 32);
}

// reads packed 2 bits per base, 32 bases per u64 (base j of a word at bits 2j), one zero pad word per read
__device__ __forceinline__ uint64_t read_key(const uint64_t *pw, int a, int cnt)
{ // the cnt (<= 16) bases starting at read position a
	uint32_t w = (uint32_t)a >> 5, sh = ((uint32_t)a & 31) * 2;
	uint64_t v = __ldg(pw + w) >> sh;
	if (sh) v |= __ldg(pw + w + 1) << (64 - sh);
	return v & ((1ull << (2 * cnt)) - 1);
}
// 1 bit per base: set where the base is ambiguous (> 3) or past the end of the read
__device__ __forceinline__ bool read_has_n(const uint32_t *pn, int a, int cnt)
{
	uint32_t w = (uint32_t)a >> 5, sh = (uint32_t)a & 31;
	uint32_t m = __ldg(pn + w) >> sh;
	if (sh) m |= __ldg(pn + w + 1) << (32 - sh);
	return (m & ((cnt >= 32 ? 0u : (1u << cnt)) - 1u)) != 0;
}

// 32 consecutive bases of a 2-bit packed sequence (base j of a word at bits 2j) starting at position pos
__device__ __forceinline__ uint64_t packed_window(const uint64_t *p, uint64_t pos)
{
	uint64_t w = pos >> 5; uint32_t sh = ((uint32_t)pos & 31) * 2;
	uint64_t v = gather_u64(p + w) >> sh;
	if (sh) v |= gather_u64(p + w + 1) << (64 - sh);
	return v;
}

// packed interval-list entry (16 bytes): three 37-bit coordinates + 16-bit read position
__device__ __forceinline__ uint4 pack_entry(uint64_t x0, uint64_t x1, uint64_t x2, uint32_t end)
{
	uint4 v;
	v.x = (uint32_t)x0; v.y = (uint32_t)x1; v.z = (uint32_t)x2;
	v.w = ((uint32_t)(x0 >> 32) & 31) | (((uint32_t)(x1 >> 32) & 31) << 5) | (((uint32_t)(x2 >> 32) & 31) << 10) | (end << 16);
	return v;
}
__device__ __forceinline__ void unpack_entry(uint4 v, uint64_t &x0, uint64_t &x1, uint64_t &x2, uint32_t &end)
{
	x0 = (uint64_t)v.x | ((uint64_t)(v.w & 31) << 32);
	x1 = (uint64_t)v.y | ((uint64_t)((v.w >> 5) & 31) << 32);
	x2 = (uint64_t)v.z | ((uint64_t)((v.w >> 10) & 31) << 32);
	end = v.w >> 16;
}
