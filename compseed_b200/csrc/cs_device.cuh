// Device-side FM-index primitives for sm_100a.
//
// HBM layout of the index (built once at upload by k_relayout from the reference's 64-byte
// buckets, FM_index/bwt.h:74-80, index_main.c:152-174):
//
//   bucket b (32 bytes, one L2/DRAM sector, 64 BWT rows [64b, 64b+64)):
//     u64 w0, w1   bases LSB-first: row 64b+i is at bits 2*(i&31) of w[i>>5]
//     u32 c0, c1, c2   low 32 bits of the number of A / C / G in rows [0, 64b)
//     u32 hi           bits 0-7 / 8-15 / 16-23: bits 32-39 of c0 / c1 / c2
//   The T checkpoint is implied: c3 = 64b - c0 - c1 - c2 (the '$' row is not stored, as in the
//   reference).  One occ4 lookup == one 256-bit load (LDG.E.256), i.e. exactly one 32-byte sector.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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
	// match that cannot grow to min_seed_len bases (see ST_PRUNE in k_seed).  0 = no filter.
	const uint32_t *pt;
	uint32_t pt_k;
};

struct Bucket { uint64_t w0, w1; uint32_t c0, c1, c2, hi; };

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
	uint64_t c01, c2h;
	const uint4 *p = I.buckets + 2 * b;
	asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
	             : "=l"(r.w0), "=l"(r.w1), "=l"(c01), "=l"(c2h) : "l"(p));
	r.c0 = (uint32_t)c01; r.c1 = (uint32_t)(c01 >> 32);
	r.c2 = (uint32_t)c2h; r.hi = (uint32_t)(c2h >> 32);
	return r;
}

// cnt[c] = number of base c in rows [0 .. 64b + r] of the stored BWT (r in 0..63, inclusive)
__device__ __forceinline__ void bucket_occ4(const Bucket &B, uint64_t b, uint32_t r, uint64_t cnt[4])
{
	const uint64_t M = 0x5555555555555555ull;
	uint32_t n = r + 1;                                  // bases counted, 1..64
	uint64_t m0 = n >= 32 ? M : (((1ull << (2 * n)) - 1) & M);
	uint64_t m1 = n <= 32 ? 0ull : (n == 64 ? M : (((1ull << (2 * (n - 32))) - 1) & M));
	uint64_t e0 = B.w0 & m0, h0 = (B.w0 >> 1) & m0;      // low / high bit planes
	uint64_t e1 = B.w1 & m1, h1 = (B.w1 >> 1) & m1;
	uint32_t nT = __popcll(e0 & h0) + __popcll(e1 & h1);
	uint32_t nH = __popcll(h0) + __popcll(h1);           // G + T
	uint32_t nE = __popcll(e0) + __popcll(e1);           // C + T
	uint32_t nG = nH - nT, nC = nE - nT, nA = n - nH - nC;
	uint64_t c0 = (uint64_t)B.c0 | ((uint64_t)(B.hi & 0xff) << 32);
	uint64_t c1 = (uint64_t)B.c1 | ((uint64_t)((B.hi >> 8) & 0xff) << 32);
	uint64_t c2 = (uint64_t)B.c2 | ((uint64_t)((B.hi >> 16) & 0xff) << 32);
	uint64_t c3 = (b << 6) - c0 - c1 - c2;
	cnt[0] = c0 + nA; cnt[1] = c1 + nC; cnt[2] = c2 + nG; cnt[3] = c3 + nT;
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
// `two` receives 1 when k and l needed two different sectors (the E2 counter of SURVEY 8d, at
// this layout's 64-row granularity).
__device__ __forceinline__ void dev_extend(const DevIndex &I, uint64_t x0, uint64_t x1, uint64_t x2, int c, int is_back,
                                           uint64_t &o0, uint64_t &o1, uint64_t &o2, uint32_t &two)
{
	uint64_t a = is_back ? x0 : x1;      // x[!is_back]
	uint64_t o = is_back ? x1 : x0;      // x[is_back]
	uint64_t k = a - 1, l = k + x2;
	uint64_t kk = k - (k >= I.primary), ll = l - (l >= I.primary);
	uint64_t bk = kk >> 6, bl = ll >> 6;
	Bucket Bk = load_bucket(I, bk), Bl;
	if (bl != bk) Bl = load_bucket(I, bl); else Bl = Bk;
	two = (bl != bk);
	uint64_t tk[4], tl[4];
	bucket_occ4(Bk, bk, (uint32_t)kk & 63, tk);
	bucket_occ4(Bl, bl, (uint32_t)ll & 63, tl);
	uint64_t s3 = tl[3] - tk[3], s2 = tl[2] - tk[2], s1 = tl[1] - tk[1], s0 = tl[0] - tk[0];
	uint64_t base = o + ((a <= I.primary) && (a + x2 - 1 >= I.primary));
	// ok[3].x[is_back] = base; ok[2] = ok[3] + s3; ok[1] = ok[2] + s2; ok[0] = ok[1] + s1
	uint64_t na, no, ns;
	if (c == 3)      { na = tk[3]; ns = s3; no = base; }
	else if (c == 2) { na = tk[2]; ns = s2; no = base + s3; }
	else if (c == 1) { na = tk[1]; ns = s1; no = base + s3 + s2; }
	else             { na = tk[0]; ns = s0; no = base + s3 + s2 + s1; }
	na += l2_at(I, c) + 1;
	o0 = is_back ? na : no;
	o1 = is_back ? no : na;
	o2 = ns;
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
	uint64_t w = r < 32 ? B.w0 : B.w1;
	uint32_t c = (uint32_t)(w >> (2 * (r & 31))) & 3;
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
	return sa + __ldg(I.sa + (k >> I.sa_shift));
}

// ---- top-of-search table and 2-bit packed reads ----
__device__ __forceinline__ uint64_t kt_offset(uint32_t d) { return ((1ull << (2 * d)) - 4) / 3; }

__device__ __forceinline__ void kt_lookup(const DevIndex &I, uint32_t d, uint64_t key, uint64_t &x0, uint64_t &x1, uint64_t &x2)
{
	uint4 v = __ldg(I.kt + kt_offset(d) + key);
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
