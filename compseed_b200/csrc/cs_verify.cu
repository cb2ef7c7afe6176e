// Self-check of a device index by the DEFINITIONS of its parts (cs_index_verify), and the random-sector probe
// over the index's own arrays (cs_probe_index_gather).  Neither is on the seeding path.
//
// Why a verifier: the headline configuration (3.1 Gbp, 6.2 G rows) is built on the GPU by cs_index_build because
// bwaidx (FM_index/index_main.c:257-325) needs hours for it, so "equal to bwaidx" can only be tested on small
// references.  What can be checked at any size is that the arrays ARE an FM-index of the given text:
//   - SA is a permutation of 0..seq_len and consecutive rows are in increasing suffix order  =>  SA is the suffix array
//     (what bwt_cal_sa / the suffix sort produce, bwt.c:62-84);
//   - BWT[r] == T[SA[r] - 1] and primary == the row with SA == 0                              =>  the BWT (bwt_gen.c);
//   - every Occ checkpoint == the running base counts, totals == L2                           =>  bwt_bwtupdate_core,
//     index_main.c:152-174;
// and that the result-neutral structures are what they claim to be: SA[ISA[p]] == p, filter count == min(3, number
// of occurrences found by backward search), table entry == the chain of bwt_extend calls from bwt_set_intv.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "cs_internal.h"

namespace {

__device__ __forceinline__ uint64_t sa_at(const DevIndex &I, uint64_t r) { return r == 0 ? I.seq_len : I.sa[r]; }   // sa[0] is the -1 sentinel (bwt.c:83)

__device__ __forceinline__ uint32_t text_base(const DevIndex &I, uint64_t p) { return (uint32_t)(I.text[p >> 5] >> (2 * (p & 31))) & 3u; }

__device__ __forceinline__ uint64_t text_window(const DevIndex &I, uint64_t pos)
{ // 32 bases from pos (base j at bits 2j); words past the text are zero-padded by the builder
	const uint64_t w = pos >> 5; const uint32_t sh = ((uint32_t)pos & 31) * 2;
	uint64_t v = I.text[w] >> sh;
	if (sh) v |= I.text[w + 1] << (64 - sh);
	return v;
}

// rows r = first, first + stride, ...: suffix order against the previous row, the BWT character, the permutation bitmap
__global__ void k_verify_rows(DevIndex I, uint64_t stride, unsigned long long *bitmap, unsigned long long *out)
{
	unsigned long long n_ord = 0, bad_ord = 0, n_bwt = 0, bad_bwt = 0, bad_perm = 0;
	const uint64_t L = I.seq_len;
	for (uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * stride; r <= L; r += (uint64_t)gridDim.x * blockDim.x * stride) {
		const uint64_t q = sa_at(I, r);
		if (q > L) { ++bad_perm; continue; }
		if (bitmap) { // every text position (and L, the '$' suffix) exactly once
			const unsigned long long bit = 1ull << (q & 63);
			if (atomicOr(bitmap + (q >> 6), bit) & bit) ++bad_perm;
		}
		// BWT character of row r (bwt_B0 with the primary adjustment, bwt.h:80, bwt.c:55-58)
		++n_bwt;
		if (r == I.primary) { if (q != 0) ++bad_bwt; }
		else if (q == 0) ++bad_bwt;               // only the primary row may hold the whole text
		else {
			const uint64_t x = r - (r > I.primary);
			const Bucket B = load_bucket(I, x >> 6);
			if (bucket_base(B, (uint32_t)x & 63) != text_base(I, q - 1)) ++bad_bwt;
		}
		if (r == 0) { if (q != L) ++bad_ord; continue; }   // row 0 is the '$' suffix
		// suffix(p) < suffix(q), '$' (the end of the text) smaller than every base
		uint64_t p = sa_at(I, r - 1), qq = q;
		++n_ord;
		if (p > L) { ++bad_ord; continue; }
		bool ok = false, decided = false;
		while (!decided) {
			const uint64_t lp = L - p, lq = L - qq;
			uint32_t nb = 32;
			if (lp < nb) nb = (uint32_t)lp;
			if (lq < nb) nb = (uint32_t)lq;
			uint64_t diff = 0, wp = 0, wq = 0;
			if (nb) {
				wp = text_window(I, p); wq = text_window(I, qq);
				diff = wp ^ wq;
				if (nb < 32) diff &= (1ull << (2 * nb)) - 1;
			}
			if (diff) {
				const uint32_t j = (uint32_t)(__ffsll((long long)diff) - 1) >> 1;
				ok = ((wp >> (2 * j)) & 3) < ((wq >> (2 * j)) & 3); decided = true;
			} else if (nb < 32) { // one of them ended: the shorter suffix is the smaller one; equal lengths would mean p == q
				ok = lp < lq; decided = true;
			} else { p += 32; qq += 32; }
		}
		if (!ok) ++bad_ord;
	}
	if (n_ord) atomicAdd(out + 0, n_ord);
	if (bad_ord) atomicAdd(out + 1, bad_ord);
	if (n_bwt) atomicAdd(out + 2, n_bwt);
	if (bad_bwt) atomicAdd(out + 3, bad_bwt);
	if (bad_perm) atomicAdd(out + 4, bad_perm);
}

// after a full pass of k_verify_rows: every bit 0..L of the bitmap must be set
__global__ void k_verify_bitmap(const unsigned long long *bitmap, uint64_t L, unsigned long long *out)
{
	unsigned long long bad = 0;
	const uint64_t n_words = (L >> 6) + 1;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long want = ~0ull;
		if (w == n_words - 1) { const uint32_t nb = (uint32_t)(L & 63) + 1; want = nb == 64 ? ~0ull : ((1ull << nb) - 1); }
		bad += (unsigned long long)__popcll(want & ~bitmap[w]);
	}
	if (bad) atomicAdd(out + 4, bad);
}

// one thread per Occ bucket: checkpoint of the next bucket == this checkpoint + the bases of this bucket; totals == L2
__global__ void k_verify_occ(DevIndex I, unsigned long long *out)
{
	unsigned long long bad = 0;
	const uint64_t nb = (I.seq_len + 63) >> 6;
	for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x) {
		const Bucket B = load_bucket(I, b);
		const uint64_t rows = I.seq_len - (b << 6) < 64 ? I.seq_len - (b << 6) : 64;
		uint64_t cnt[4];
		bucket_occ4(B, b, (uint32_t)rows - 1, cnt);                      // counts over stored rows [0, 64b + rows)
		const uint64_t e1 = cnt[0], e2 = cnt[0] + cnt[1], e3 = e2 + cnt[2];
		if (b == 0 && (bucket_p(B, 1) | bucket_p(B, 2) | bucket_p(B, 3))) ++bad;
		if (cnt[0] + cnt[1] + cnt[2] + cnt[3] != (b << 6) + rows) ++bad;
		if (b + 1 < nb) {
			const Bucket N = load_bucket(I, b + 1);
			if (bucket_p(N, 1) != e1 || bucket_p(N, 2) != e2 || bucket_p(N, 3) != e3) ++bad;
		} else if (e1 != I.L2[1] - I.L2[0] || e2 != I.L2[2] - I.L2[0] || e3 != I.L2[3] - I.L2[0] || I.L2[4] != I.seq_len || I.L2[0] != 0) ++bad;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0 && I.primary > I.seq_len) ++bad;
	if (bad) atomicAdd(out + 5, bad);
}

__global__ void k_verify_isa(DevIndex I, uint64_t stride, unsigned long long *out)
{
	unsigned long long n = 0, bad = 0;
	const uint64_t n_isa = (I.seq_len >> I.isa_shift) + 1;               // samples at p = 0, 2^shift, ... <= seq_len
	for (uint64_t j = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * stride; j < n_isa; j += (uint64_t)gridDim.x * blockDim.x * stride) {
		const uint64_t row = I.isa[j];
		++n;
		if (row > I.seq_len || sa_at(I, row) != (j << I.isa_shift)) ++bad;
	}
	if (n) atomicAdd(out + 6, n);
	if (bad) atomicAdd(out + 7, bad);
}

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{ x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

// sample i even: the K-mer at a pseudo-random text position; odd: a pseudo-random K-mer.  Occurrences by backward search
// with the same dev_extend the seeding kernels use; the filter must say min(3, occurrences).
__global__ void k_verify_filter(DevIndex I, uint64_t n_samples, unsigned long long *out)
{
	unsigned long long n = 0, bad = 0;
	const uint32_t K = I.pt_k;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_samples; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t key;
		const uint64_t h = mix64(i * 0x9E3779B97F4A7C15ull + 12345);
		if (i & 1) key = h & ((1ull << (2 * K)) - 1);
		else {
			const uint64_t p = (uint64_t)(((unsigned __int128)h * (I.seq_len - K + 1)) >> 64);
			key = text_window(I, p) & ((1ull << (2 * K)) - 1);
		}
		// backward search: last base first (bwt_set_intv, bwt.h:82), then prepend
		int b = (int)((key >> (2 * (K - 1))) & 3);
		uint64_t x0 = l2_at(I, b) + 1, x1 = l2_at(I, 3 - b) + 1, x2 = l2_at(I, b + 1) - l2_at(I, b);
		for (int j = (int)K - 2; j >= 0 && x2 > 0; --j) {
			uint64_t o0, o1, o2; uint32_t two;
			dev_extend(I, x0, x1, x2, (int)((key >> (2 * j)) & 3), 1, o0, o1, o2, two);
			x0 = o0; x1 = o1; x2 = o2;
		}
		const uint32_t want = x2 > 3 ? 3u : (uint32_t)x2;
		const uint32_t got = (I.pt[key >> 4] >> (2 * ((uint32_t)key & 15))) & 3;
		++n;
		if (got != want) ++bad;
	}
	if (n) atomicAdd(out + 8, n);
	if (bad) atomicAdd(out + 9, bad);
}

// sampled table entries of every depth against the chain of forward bwt_extend calls from the one-base interval
__global__ void k_verify_table(DevIndex I, uint64_t n_samples, unsigned long long *out)
{
	unsigned long long n = 0, bad = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_samples; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t d = 1 + (uint32_t)(i % I.kt_depth);
		uint64_t key;
		const uint64_t h = mix64(i * 0xD1B54A32D192ED03ull + 777);
		if ((i / I.kt_depth) & 1) key = h & ((1ull << (2 * d)) - 1);
		else {
			const uint64_t p = (uint64_t)(((unsigned __int128)h * (I.seq_len - d + 1)) >> 64);
			key = text_window(I, p) & ((1ull << (2 * d)) - 1);
		}
		int b = (int)(key & 3);
		uint64_t x0 = l2_at(I, b) + 1, x1 = l2_at(I, 3 - b) + 1, x2 = l2_at(I, b + 1) - l2_at(I, b);
		for (uint32_t j = 1; j < d && x2 > 0; ++j) { // appending base c == prepending its complement on the other strand (bwt.c:309)
			uint64_t o0, o1, o2; uint32_t two;
			dev_extend(I, x0, x1, x2, 3 - (int)((key >> (2 * j)) & 3), 0, o0, o1, o2, two);
			x0 = o0; x1 = o1; x2 = o2;
		}
		uint64_t t0, t1, t2;
		kt_lookup(I, d, key, t0, t1, t2);
		++n;
		if (x2 == 0 ? t2 != 0 : (t0 != x0 || t1 != x1 || t2 != x2)) ++bad;
	}
	if (n) atomicAdd(out + 10, n);
	if (bad) atomicAdd(out + 11, bad);
}

// repeat lengths by their definition, with the FM-index instead of the suffix array's neighbours: T[p, p+R) must have at
// least two occurrences (backward search over its R bases), T[p, p+R+1) exactly one (unless R is the cap or the text ends)
__global__ void k_verify_rep(DevIndex I, uint64_t n_samples, unsigned long long *out)
{
	unsigned long long n = 0, bad = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_samples; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t h = mix64(i * 0xA24BAED4963EE407ull + 4242);
		const uint64_t p = (uint64_t)(((unsigned __int128)h * I.seq_len) >> 64);
		const uint32_t R = I.rep[p];
		++n;
		if (p + R > I.seq_len) { ++bad; continue; }
		const bool more = R < 255 && p + R < I.seq_len;              // the string one base longer exists and is not capped
		const uint32_t len = R + (more ? 1u : 0u);
		if (len == 0) continue;
		// backward search over T[p, p+len): occurrences of every suffix of it; the count of T[p, p+len) is the last one,
		// that of T[p+1, ...) irrelevant -- so search the REVERSED roles: forward extension from T[p] (bwt.c:309)
		int b = (int)text_base(I, p);
		uint64_t x0 = l2_at(I, b) + 1, x1 = l2_at(I, 3 - b) + 1, x2 = l2_at(I, b + 1) - l2_at(I, b), at_R = R == 1 ? x2 : 0;
		for (uint32_t j = 1; j < len && x2 > 0; ++j) {
			uint64_t o0, o1, o2; uint32_t two;
			dev_extend(I, x0, x1, x2, 3 - (int)text_base(I, p + j), 0, o0, o1, o2, two);
			x0 = o0; x1 = o1; x2 = o2;
			if (j + 1 == R) at_R = x2;
		}
		if (R >= 1 && at_R < 2) ++bad;
		else if (more && x2 != 1) ++bad;
	}
	if (n) atomicAdd(out + 14, n);
	if (bad) atomicAdd(out + 15, bad);
}

// a chunk of the forward strand as the caller holds it (nt4 bytes): T[p] == fwd[p], T[2 l_pac - 1 - p] == 3 - fwd[p]
__global__ void k_verify_text(DevIndex I, const uint8_t *chunk, uint64_t p0, uint64_t n_chunk, uint64_t l_pac, unsigned long long *out)
{
	unsigned long long n = 0, bad = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_chunk; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t p = p0 + i;
		const uint32_t c = chunk[i];
		n += 2;
		if (c > 3) { bad += 2; continue; }
		if (text_base(I, p) != c) ++bad;
		if (text_base(I, 2 * l_pac - 1 - p) != 3 - c) ++bad;
	}
	if (n) atomicAdd(out + 12, n);
	if (bad) atomicAdd(out + 13, bad);
}

// random 32-byte sectors over several arrays at once: load i picks array a with probability size_a / total
struct ProbeArrays { const uint4 *base[7]; uint64_t cum[8]; uint64_t total; int n; };   // cum: cumulative sizes in 32-byte sectors

template <int UNROLL>
__global__ void k_index_gather(ProbeArrays A, uint64_t n_loads, uint64_t seed, unsigned long long *sink)
{
	const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t s = seed ^ (tid * 0x9E3779B97F4A7C15ull);
	uint32_t acc = 0;
	const uint64_t total = A.total;           // (no run-time index into the parameter struct: that would move it to local memory)
	for (uint64_t i = tid; i < n_loads; i += stride * UNROLL) {
		uint64_t a[UNROLL], b[UNROLL], c[UNROLL], d[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			s ^= s << 13; s ^= s >> 7; s ^= s << 17;
			uint64_t g = (uint64_t)(((unsigned __int128)s * total) >> 64);
			const uint4 *bp = A.base[0]; uint64_t c0 = 0;          // (selects, not a run-time index into the parameter struct)
#pragma unroll
			for (int j = 1; j < 7; ++j) if (g >= A.cum[j] && A.cum[j + 1] > A.cum[j]) { bp = A.base[j]; c0 = A.cum[j]; }
			const uint4 *p = bp + 2 * (g - c0);
			asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[u]), "=l"(b[u]), "=l"(c[u]), "=l"(d[u]) : "l"(p));
		}
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) acc += (uint32_t)(a[u] ^ b[u] ^ c[u] ^ d[u]);
	}
	if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

} // namespace

extern "C" int cs_index_verify(const cs_index_t *idx, const uint8_t *fwd, uint64_t l_pac, uint32_t stride, uint64_t out[16])
{
	unsigned long long *d_out = nullptr, *d_bitmap = nullptr;
	uint8_t *d_chunk = nullptr;
	if (!idx || !out) return cs_set_err(CS_E_ARG, "null argument");
	if (stride == 0) stride = 1;
	const DevIndex &I = idx->d;
	if (I.sa_mask != 0 || !I.text) return cs_set_err(CS_E_ARG, "cs_index_verify needs the dense suffix array and the 2-bit text (sa_intv 1, isa_intv > 0)");
	if (fwd && 2 * l_pac != I.seq_len) return cs_set_err(CS_E_ARG, "fwd has %llu bases, the index %llu rows", (unsigned long long)l_pac, (unsigned long long)I.seq_len);
	{ const int rc = cs_use_device(idx->device); if (rc != CS_OK) return rc; }
	{
		const int grid = idx->n_sm * 8;
		CK(cudaMalloc(&d_out, 16 * 8));
		CK(cudaMemset(d_out, 0, 16 * 8));
		if (stride == 1) {
			const uint64_t words = (I.seq_len >> 6) + 1;
			CK(cudaMalloc(&d_bitmap, words * 8));
			CK(cudaMemset(d_bitmap, 0, words * 8));
		}
		k_verify_rows<<<grid, 256>>>(I, (uint64_t)stride, d_bitmap, d_out);
		CK(cudaGetLastError());
		if (d_bitmap) { k_verify_bitmap<<<grid, 256>>>(d_bitmap, I.seq_len, d_out); CK(cudaGetLastError()); }
		k_verify_occ<<<grid, 256>>>(I, d_out);
		CK(cudaGetLastError());
		if (I.isa) { k_verify_isa<<<grid, 256>>>(I, (uint64_t)stride, d_out); CK(cudaGetLastError()); }
		{
			const uint64_t n_samples = std::max<uint64_t>(1u << 16, std::min<uint64_t>(I.seq_len / stride, 1ull << 24));
			if (I.pt && I.pt_k >= 2 && I.seq_len > I.pt_k) { k_verify_filter<<<grid, 256>>>(I, n_samples, d_out); CK(cudaGetLastError()); }
			if (I.kt && I.kt_depth >= 1 && I.seq_len > I.kt_depth) { k_verify_table<<<grid, 256>>>(I, n_samples, d_out); CK(cudaGetLastError()); }
			if (I.rep) { k_verify_rep<<<grid, 256>>>(I, n_samples, d_out); CK(cudaGetLastError()); }
		}
		if (fwd) {
			const uint64_t chunk = 256ull << 20;
			CK(cudaMalloc(&d_chunk, (size_t)std::min<uint64_t>(chunk, l_pac)));
			for (uint64_t p0 = 0; p0 < l_pac; p0 += chunk) {
				const uint64_t nc = std::min<uint64_t>(chunk, l_pac - p0);
				CK(cudaMemcpy(d_chunk, fwd + p0, (size_t)nc, cudaMemcpyHostToDevice));
				k_verify_text<<<grid, 256>>>(I, d_chunk, p0, nc, l_pac, d_out);
				CK(cudaGetLastError());
			}
		}
		CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(out, d_out, 16 * 8, cudaMemcpyDeviceToHost));
	}
	cudaFree(d_out); cudaFree(d_bitmap); cudaFree(d_chunk);
	return CS_OK;
fail:
	cudaFree(d_out); cudaFree(d_bitmap); cudaFree(d_chunk);
	return CS_E_CUDA;
}

extern "C" int cs_probe_index_gather(const cs_index_t *idx, uint64_t n_loads, int iters, int unroll, double *gbytes_per_s, double *gloads_per_s)
{
	unsigned long long *d_sink = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	float best = 1e30f;
	ProbeArrays A;
	if (!idx) return cs_set_err(CS_E_ARG, "null index");
	{ const int rc = cs_use_device(idx->device); if (rc != CS_OK) return rc; }
	memset(&A, 0, sizeof A);
	{
		const DevIndex &I = idx->d;
		auto add = [&](const void *p, uint64_t bytes) { if (p && bytes >= 32) { A.base[A.n] = (const uint4*)p; A.cum[A.n + 1] = A.cum[A.n] + bytes / 32; ++A.n; } };
		add(I.buckets, I.n_buckets * 32);
		add(I.sa, I.n_sa * 8);
		if (I.pt) add(I.pt, ((1ull << (2 * I.pt_k)) / 16) * 4);
		if (I.text) add(I.text, ((I.seq_len + 31) / 32) * 8);
		if (I.isa) add(I.isa, ((I.seq_len >> I.isa_shift) + 1) * 8);
		if (I.kt) add(I.kt, (((1ull << (2 * (I.kt_depth + 1))) - 4) / 3) * 16);
		if (I.rep) add(I.rep, I.seq_len & ~31ull);
		for (int k = A.n; k < 7; ++k) { A.base[k] = A.base[0]; A.cum[k + 1] = A.cum[A.n]; }
		A.total = A.cum[A.n];
	}
	CK(cudaMalloc(&d_sink, 8));
	CK(cudaMemset(d_sink, 0, 8));
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int it = 0; it < iters + 1; ++it) {
		CK(cudaEventRecord(e0));
		if (unroll >= 4) k_index_gather<4><<<idx->n_sm * 8, 256>>>(A, n_loads, 0x1234567ull + it, d_sink);
		else k_index_gather<1><<<idx->n_sm * 8, 256>>>(A, n_loads, 0x1234567ull + it, d_sink);
		CK(cudaGetLastError());
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (it > 0 && ms < best) best = ms;
	}
	if (gloads_per_s) *gloads_per_s = (double)n_loads / (best * 1e-3) * 1e-9;
	if (gbytes_per_s) *gbytes_per_s = (double)n_loads * 32 / (best * 1e-3) * 1e-9;
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_sink);
	return CS_OK;
fail:
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	cudaFree(d_sink);
	return CS_E_CUDA;
}
