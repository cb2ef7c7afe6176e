// Hand-written sm_100a kernels of the SMEM seeding path.  See DESIGN.md for the layout, the
// per-kernel roofline and the mapping to the reference (FM_index/bwt.c, mapping/bwamem.c:218-272,
// mapping/comp_seed.cpp:67-160,2255-2346).
#include "cs_kernels.cuh"

// ---------------------------------------------------------------------------------------------
// Index re-layout: reference 64-byte buckets (4 x u64 checkpoint + 8 x u32, 16 bases per word,
// base i at bits (15-i)*2; index_main.c:152-174) -> 32-byte device buckets (cs_device.cuh).
// One thread per device bucket.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t msb_to_lsb16(uint32_t w)
{ // reverse the order of the sixteen 2-bit fields
	uint32_t v = __brev(w);
	return ((v & 0x55555555u) << 1) | ((v >> 1) & 0x55555555u);
}

__global__ void k_relayout(const uint32_t *src, uint64_t src_words, uint64_t seq_len, uint4 *dst, uint64_t n_buckets)
{
	for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_buckets; b += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t B = b >> 1, base = B << 4;       // source bucket, its first word
		uint32_t half = (uint32_t)b & 1;
		uint64_t cnt[4] = {0, 0, 0, 0};
		uint32_t w[8];
		if (b << 6 < seq_len) {
			const uint64_t *cp = reinterpret_cast<const uint64_t*>(src + base);
			cnt[0] = cp[0]; cnt[1] = cp[1]; cnt[2] = cp[2]; cnt[3] = cp[3];
		}
		for (int j = 0; j < 8; ++j) {
			uint64_t row = (B << 7) + 16 * j, idx = base + 8 + j;
			w[j] = (row < seq_len && idx < src_words) ? src[idx] : 0u;
		}
		if (half) { // add the first 64 rows of the source bucket to the checkpoint
			for (int j = 0; j < 4; ++j)
				for (int i = 0; i < 16; ++i) {
					uint64_t row = (B << 7) + 16 * j + i;
					if (row < seq_len) ++cnt[(w[j] >> ((15 - i) << 1)) & 3];
				}
		}
		const uint32_t *ws = w + 4 * half;
		uint64_t w0 = (uint64_t)msb_to_lsb16(ws[0]) | ((uint64_t)msb_to_lsb16(ws[1]) << 32);
		uint64_t w1 = (uint64_t)msb_to_lsb16(ws[2]) | ((uint64_t)msb_to_lsb16(ws[3]) << 32);
		uint4 lo, hi;
		lo.x = (uint32_t)w0; lo.y = (uint32_t)(w0 >> 32); lo.z = (uint32_t)w1; lo.w = (uint32_t)(w1 >> 32);
		const uint64_t p1 = cnt[0], p2 = p1 + cnt[1], p3 = p2 + cnt[2];   // prefix-sum checkpoint (cs_device.cuh)
		hi.x = (uint32_t)p1; hi.y = (uint32_t)p2; hi.z = (uint32_t)p3;
		hi.w = ((uint32_t)(p1 >> 32) & 0xff) | (((uint32_t)(p2 >> 32) & 0xff) << 8) | (((uint32_t)(p3 >> 32) & 0xff) << 16);
		dst[2 * b] = lo; dst[2 * b + 1] = hi;
	}
}

// Inverse of k_relayout: rebuild the reference layout (for cs_index_download).  One thread per
// reference bucket (128 rows); the trailing checkpoint-only record is written by the last thread.
__global__ void k_unlayout(const uint4 *src, uint64_t seq_len, uint32_t *dst, uint64_t dst_words)
{
	uint64_t n_ref = (seq_len + 127) >> 7;
	for (uint64_t B = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; B <= n_ref; B += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t row0 = B << 7;
		// checkpoint at row0: from device bucket 2B if it exists, else from the last bucket's totals
		uint64_t cnt[4];
		uint64_t b = row0 >> 6;
		uint64_t nb = (seq_len + 63) >> 6;
		if (b < nb) {
			uint4 h = src[2 * b + 1];
			uint64_t p1 = (uint64_t)h.x | ((uint64_t)(h.w & 0xff) << 32);
			uint64_t p2 = (uint64_t)h.y | ((uint64_t)((h.w >> 8) & 0xff) << 32);
			uint64_t p3 = (uint64_t)h.z | ((uint64_t)((h.w >> 16) & 0xff) << 32);
			cnt[0] = p1; cnt[1] = p2 - p1; cnt[2] = p3 - p2; cnt[3] = row0 - p3;
		} else { // row0 >= seq_len: totals = checkpoint of the last bucket + its bases
			uint64_t lb = nb - 1;
			uint4 l = src[2 * lb], h = src[2 * lb + 1];
			uint64_t p1 = (uint64_t)h.x | ((uint64_t)(h.w & 0xff) << 32);
			uint64_t p2 = (uint64_t)h.y | ((uint64_t)((h.w >> 8) & 0xff) << 32);
			uint64_t p3 = (uint64_t)h.z | ((uint64_t)((h.w >> 16) & 0xff) << 32);
			cnt[0] = p1; cnt[1] = p2 - p1; cnt[2] = p3 - p2; cnt[3] = (lb << 6) - p3;
			uint64_t w0 = (uint64_t)l.x | ((uint64_t)l.y << 32), w1 = (uint64_t)l.z | ((uint64_t)l.w << 32);
			for (uint64_t r = lb << 6; r < seq_len; ++r) {
				uint32_t i = (uint32_t)(r & 63);
				++cnt[((i < 32 ? w0 : w1) >> (2 * (i & 31))) & 3];
			}
		}
		// position of this record in the reference array: 16 words per full bucket
		uint64_t base = B << 4;
		if (B == n_ref) { // trailing record sits right after the last BWT word (index_main.c:171)
			base = ((n_ref - (n_ref > 0)) << 4);
			if (n_ref > 0) base += 8 + (((seq_len - ((n_ref - 1) << 7)) + 15) >> 4);
		}
		if (base + 8 <= dst_words) { // the trailing record may be only 4-byte aligned: write 32-bit halves
			for (int c = 0; c < 4; ++c) { dst[base + 2 * c] = (uint32_t)cnt[c]; dst[base + 2 * c + 1] = (uint32_t)(cnt[c] >> 32); }
		}
		if (B < n_ref) {
			for (int j = 0; j < 8; ++j) {
				uint64_t row = row0 + 16 * j;
				if (row >= seq_len) break;
				uint64_t bb = row >> 6;
				uint4 l = src[2 * bb];
				uint32_t q = (uint32_t)(row & 63) >> 4;   // which 16-base quarter of the device bucket
				uint32_t v = q == 0 ? l.x : q == 1 ? l.y : q == 2 ? l.z : l.w;
				uint32_t r = ((v & 0x55555555u) << 1) | ((v >> 1) & 0x55555555u);
				dst[base + 8 + j] = __brev(r);
			}
		}
	}
}

// Re-sample the suffix array at a denser (or sparser) interval: out[r >> out_shift] = bwt_sa(r).
__global__ void k_resample_sa(DevIndex I, uint64_t *out, uint64_t n_out, uint32_t out_shift)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t steps;
		out[i] = i == 0 ? (uint64_t)-1 : dev_sa(I, i << out_shift, steps);
	}
}

// ---------------------------------------------------------------------------------------------
// Unit-level probes (bwt_occ4 / bwt_extend twins)
// ---------------------------------------------------------------------------------------------
__global__ void k_probe_occ4(DevIndex I, uint32_t n, const uint64_t *k, uint64_t *cnt)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t c[4] = {0, 0, 0, 0};
	if (k[i] != (uint64_t)-1) dev_occ4(I, k[i], c);
	for (int j = 0; j < 4; ++j) cnt[4 * (size_t)i + j] = c[j];
}

__global__ void k_probe_extend(DevIndex I, uint32_t n, const uint64_t *ik, const int32_t *is_back, uint64_t *ok)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t in[3] = { ik[3 * (size_t)i], ik[3 * (size_t)i + 1], ik[3 * (size_t)i + 2] };
	uint64_t out4[12];
	dev_extend4(I, in, is_back[i] != 0, out4);
	// cross-check the single-child path used by the seeding kernel against the 4-child path
	if (in[(is_back[i] != 0) ? 0 : 1] >= 1 && in[2] > 0) {
		for (int c = 0; c < 4; ++c) {
			uint64_t o0, o1, o2; uint32_t two;
			dev_extend(I, in[0], in[1], in[2], c, is_back[i] != 0, o0, o1, o2, two);
			if (o0 != out4[c * 3] || o1 != out4[c * 3 + 1] || o2 != out4[c * 3 + 2]) out4[c * 3 + 2] = ~0ull; // poison
		}
	}
	for (int j = 0; j < 12; ++j) ok[12 * (size_t)i + j] = out4[j];
}

// ---------------------------------------------------------------------------------------------
// Top-of-search table: depth d from depth d-1, one thread per parent string (4 children each).
// ---------------------------------------------------------------------------------------------
__global__ void k_kt_build(DevIndex I, uint4 *kt, uint32_t d)
{
	const uint64_t n_parent = 1ull << (2 * (d - 1));
	for (uint64_t pk = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; pk < n_parent; pk += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t ok[12];
		if (d == 1) { // bwt_set_intv, bwt.h:82; stored in the order dev_extend4 uses: child of base b at index 3-b
			for (int b = 0; b < 4; ++b) { ok[(3 - b) * 3] = I.L2[b] + 1; ok[(3 - b) * 3 + 1] = I.L2[3 - b] + 1; ok[(3 - b) * 3 + 2] = I.L2[b + 1] - I.L2[b]; }
		} else {
			uint4 v = kt[kt_offset(d - 1) + pk];
			uint64_t ik[3];
			ik[0] = (uint64_t)v.x | ((uint64_t)(v.w & 31) << 32);
			ik[1] = (uint64_t)v.y | ((uint64_t)((v.w >> 5) & 31) << 32);
			ik[2] = (uint64_t)v.z | ((uint64_t)((v.w >> 10) & 31) << 32);
			if (ik[2] == 0) { for (int j = 0; j < 12; ++j) ok[j] = 0; }
			else dev_extend4(I, ik, 0, ok); // appending base b == prepending its complement on the reverse strand (bwt.c:309)
		}
		for (int b = 0; b < 4; ++b) {
			const uint64_t *c = ok + (3 - b) * 3;
			uint64_t x2 = c[2];
			kt[kt_offset(d) + (pk | ((uint64_t)b << (2 * (d - 1))))] = x2 ? pack_entry(c[0], c[1], x2, 0) : make_uint4(0, 0, 0, 0);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Occurrence filter construction.
// k_text_from_index: rebuild the indexed text T (2 bits/base, 32 bases per u64, base j of a word at
// bits 62-2j) from the BWT and the sampled SA: one LF walk per sampled row, each writing the
// characters between two samples (T[SA[r]-1] is the BWT character of row r).
// k_pt_count: saturating 2-bit count of every K-mer of T.
// ---------------------------------------------------------------------------------------------
__global__ void k_text_from_index(DevIndex I, unsigned long long *W)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < I.n_sa; s += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t r = s << I.sa_shift;
		uint64_t p = s == 0 ? I.seq_len : I.sa[s];
		for (;;) {
			if (r == I.primary || p == 0) break;              // the '$' row: nothing precedes text position 0
			uint64_t x = r - (r > I.primary);
			uint64_t b = x >> 6; uint32_t q = (uint32_t)x & 63;
			Bucket B = load_bucket(I, b);
			uint32_t c = bucket_base(B, q);
			--p;
			if (c) atomicOr(W + (p >> 5), (unsigned long long)c << (62 - 2 * (p & 31)));
			uint64_t cnt[4];
			bucket_occ4(B, b, q, cnt);
			r = l2_at(I, (int)c) + (c == 0 ? cnt[0] : c == 1 ? cnt[1] : c == 2 ? cnt[2] : cnt[3]);
			if ((r & I.sa_mask) == 0) break;                  // the next sample's walk takes over
		}
	}
}

// text in the bit order of the packed reads (base j of a word at bits 2j) from the builder's order (bits 62-2j)
__global__ void k_text_lsb(const uint64_t *W, uint64_t n_words, uint64_t *out)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t v = __brevll(W[i]);                     // reverses bits: base order fixed, bit order inside a base swapped
		out[i] = ((v & 0x5555555555555555ull) << 1) | ((v >> 1) & 0x5555555555555555ull);
	}
}

// inverse suffix array sampled every 2^shift text positions, from the dense SA (sa[r] = text position of row r)
__global__ void k_isa_sample(DevIndex I, uint64_t *isa, uint32_t shift)
{
	const uint64_t mask = (1ull << shift) - 1;
	for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < I.n_sa; r += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t p = r == 0 ? I.seq_len : I.sa[r];
		if ((p & mask) == 0) isa[p >> shift] = r;
	}
}

// Repeat lengths (DevIndex::rep) from the dense SA and the 2-bit text: suffix r shares max(lcp(r-1, r), lcp(r, r+1)) bases with
// some other suffix, and no more.
__device__ __forceinline__ uint32_t text_lcp(const DevIndex &I, uint64_t p, uint64_t q)   // common prefix of the suffixes p and q of T, capped at 255
{
	uint64_t room = I.seq_len - (p > q ? p : q);
	if (room > 255) room = 255;
	uint32_t n = 0;
	while (n < room) {
		const uint64_t diff = packed_window(I.text, p + n) ^ packed_window(I.text, q + n);
		const uint32_t m = diff ? (uint32_t)(__ffsll((long long)diff) - 1) >> 1 : 32u;
		n += m;
		if (m < 32) break;
	}
	return n < room ? n : (uint32_t)room;
}
__global__ void k_rep_build(DevIndex I, uint8_t *rep)
{
	for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1; r <= I.seq_len; r += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t p = I.sa[r];
		const uint32_t a = r > 1 ? text_lcp(I, p, I.sa[r - 1]) : 0u;
		const uint32_t b = r < I.seq_len ? text_lcp(I, p, I.sa[r + 1]) : 0u;
		rep[p] = (uint8_t)(a > b ? a : b);
	}
}

#define PT_CHUNK 256
__global__ void k_pt_count(const uint64_t *W, uint64_t n, uint32_t K, uint32_t *pt)
{
	if (n < K) return;
	const uint64_t n_pos = n - K + 1;
	for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * PT_CHUNK; i0 < n_pos; i0 += (uint64_t)gridDim.x * blockDim.x * PT_CHUNK) {
		uint64_t key = 0;
		for (uint32_t j = 0; j + 1 < K; ++j) { uint64_t p = i0 + j; key |= ((W[p >> 5] >> (62 - 2 * (p & 31))) & 3) << (2 * (j + 1)); }
		uint64_t i1 = i0 + PT_CHUNK < n_pos ? i0 + PT_CHUNK : n_pos;
		for (uint64_t i = i0; i < i1; ++i) { // key of T[i..i+K): base j at bits 2j
			uint64_t p = i + K - 1;
			key = (key >> 2) | (((W[p >> 5] >> (62 - 2 * (p & 31))) & 3) << (2 * (K - 1)));
			uint32_t *w = pt + (key >> 4); uint32_t sh = 2 * ((uint32_t)key & 15);
			uint32_t old = *w;
			while (((old >> sh) & 3) != 3) {
				uint32_t seen = atomicCAS(w, old, old + (1u << sh));
				if (seen == old) break;
				old = seen;
			}
		}
	}
}

// 2-bit packing of the reads of a batch + ambiguity mask.  Eight lanes per read (four reads per warp), one lane
// per 32-base word; a word is built from nine aligned 32-bit loads (the read may start at any byte), four
// bases per SIMD step.  The byte buffer is padded so that the aligned loads past a read's end stay inside it.
__global__ void k_pack_reads(const uint8_t *bases, const uint32_t *off, uint32_t n_reads, uint64_t *packed, uint32_t *nmask)
{
	const uint32_t sub = threadIdx.x & 7;
	const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
	for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n_reads; r += ngroups) {
		const uint32_t o = off[r], len = off[r + 1] - o;
		const uint64_t w0 = (uint64_t)(o >> 5) + 2 * r;
		const uint32_t nw = (len >> 5) + 2;
		for (uint32_t w = sub; w < nw; w += 8) {
			uint64_t v = 0; uint32_t m = 0xffffffffu;
			const uint32_t p0 = w << 5;
			if (p0 < len) {
				const uint64_t addr = (uint64_t)o + p0;
				const uint32_t *ap = reinterpret_cast<const uint32_t*>(bases + (addr & ~3ull));
				const uint32_t sh = ((uint32_t)addr & 3) * 8;
				uint32_t u = __ldg(ap);
				m = 0;
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const uint32_t un = __ldg(ap + k + 1);
					const uint32_t x = __funnelshift_r(u, un, sh);              // bases p0+4k .. p0+4k+3, one per byte
					u = un;
					const uint32_t amb = __vcmpgtu4(x, 0x03030303u);            // 0xff where the code is > 3
					uint32_t c = x & 0x03030303u & ~amb;
					c |= c >> 6; c = (c | (c >> 12)) & 0xffu;                    // four 2-bit codes, base j at bits 2j
					uint32_t a1 = amb & 0x01010101u;
					a1 = (a1 | (a1 >> 7) | (a1 >> 14) | (a1 >> 21)) & 0xfu;
					v |= (uint64_t)c << (8 * k);
					m |= a1 << (4 * k);
				}
				const uint32_t valid = len - p0;                                // bases of this word inside the read
				if (valid < 32) { v &= (1ull << (2 * valid)) - 1; m |= ~((1u << valid) - 1u); }
			}
			packed[w0 + w] = v; nmask[w0 + w] = m;
		}
	}
}

// Inverse of k_pack_reads, for batches submitted in packed form (cs_seed_batch_submit_packed) when a kernel that
// reads the byte form is going to run (k_seed_r3).  One thread per base.
__global__ void k_unpack_reads(const uint64_t *packed, const uint32_t *nmask, const uint32_t *off, uint32_t n_reads, uint8_t *bases, uint32_t off_bias)
{
	const uint32_t sub = threadIdx.x & 7;
	const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
	for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n_reads; r += ngroups) {
		const uint32_t o = off[r], len = off[r + 1] - o;
		const uint64_t w0 = (uint64_t)((o + off_bias) >> 5) + 2 * r;
		for (uint32_t p = sub; p < len; p += 8) {
			const uint64_t v = packed[w0 + (p >> 5)];
			const uint32_t m = nmask[w0 + (p >> 5)];
			bases[o + p] = ((m >> (p & 31)) & 1) ? 4 : (uint8_t)((v >> (2 * (p & 31))) & 3);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// SA resolution: rows_inout[i] = bwt_sa(rows_inout[i]).  LF walks have a geometric length
// distribution (mean 31, max 396+ at sa_intv 32), so lanes refill from a global
// work counter as soon as their own walk ends instead of waiting for the slowest lane.
// ---------------------------------------------------------------------------------------------
__global__ void k_sa_resolve(DevIndex I, const uint32_t *n_ptr, uint64_t cap, uint64_t *rows_inout, unsigned long long *work,
                             unsigned long long *lf_steps)
{
	unsigned long long steps_total = 0;
	uint64_t n = *n_ptr;           // produced on the device by the collect pass: no host round trip
	if (n > cap) return;           // overflow is reported by the collect pass
	if (I.sa_mask == 0) { // dense suffix array: one 8-byte gather per seed
		for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
			rows_inout[i] = gather_u64(I.sa + rows_inout[i]);
		return;
	}
	for (;;) {
		unsigned long long i = atomicAdd(work, 1ull);
		if (i >= n) break;
		uint32_t steps;
		rows_inout[i] = dev_sa(I, rows_inout[i], steps);
		steps_total += steps;
	}
	if (steps_total) atomicAdd(lf_steps, steps_total);
}

// ---------------------------------------------------------------------------------------------
// The LITERAL seeding kernel (passes 1 and 2 of mem_collect_intv): bwt_smem1a exactly as written, interval lists
// and all.  Two modes: read mode (a.defer_q == NULL: every read of the batch; used when k_seed_fast is not
// applicable -- sampled SA, long reads, min_seed_len below the filter's K) and call mode (the bwt_smem1a calls
// k_seed_fast / k_seed_walk could not prove simple, one per lane, plus the second-pass calls of what they find).
//
// One read per THREAD, persistent threads, reads handed out in input order by an atomic counter
// (neighbouring reordered reads run at the same time on neighbouring threads, so the sectors they
// share are hit in L1/L2).  Every thread runs bwt_smem1a as an explicit state machine whose only
// expensive step -- one bwt_extend == one or two 32-byte sector gathers -- is executed convergently
// by the whole warp once per trip of the outer loop (__all_sync re-converges the 32 lanes every
// trip; idle lanes never leave early); the cheap, divergent bookkeeping (list pushes, mem emission,
// pivot selection) happens in between.  With ~114k threads resident this keeps ~150k independent
// random sector reads in flight, which is what a latency-bound dependent-gather workload needs
// (hard part 3).
//
// Shared memory per thread: the interval list of bwt_smem1a (prev/curr, FM_index/bwt.c:293-344) as
// ONE in-place stack of packed 16-byte entries ([entry][thread], bank-conflict free; entries beyond
// CS_LIST_SMEM spill to HBM), and the read in flight, 2 bits per base + N mask (RW 32-base words;
// RW == 0: read from global memory, for reads longer than 32*RW bases).
//
// Occurrence filter (result-neutral): at the start of every bwt_smem1a
// call the warp tests, for one requesting lane at a time, the K-mers q[e-K, e) for e = x+1 .. x+K-1
// (one 2-bit gather per lane).  A forward match [x, e) shorter than K whose K-mer window occurs
// fewer than min_intv times, or runs into the read start / an N, cannot grow to min_seed_len >= K
// bases: its own mem would be dropped by the length filter (bwamem.c:231-233,247) and it can neither
// block nor unblock a longer match (it dies no later than any longer one), so it is never pushed.
// ---------------------------------------------------------------------------------------------
enum { ST_FETCH = 0, ST_R1_PIVOT, ST_FILTER, ST_FWD, ST_BWD_INIT, ST_BWD_SWEEP, ST_BWD_ENTRY, ST_CALL_DONE, ST_R2_NEXT,
       ST_READ_DONE, ST_TXT_SA, ST_TXT_CMP, ST_BTX_SA, ST_BTX_CMP, ST_ROW_ISA, ST_ROW_LF, ST_IDLE };
#ifndef CS_BK_QUORUM
#define CS_BK_QUORUM 1     // lanes that must be waiting before the warp runs the divergent bookkeeping section (1 = every trip)
#endif
#ifdef CS_STATS   // diagnostics build: event counters (scripts/spec_stats.py)
#define STAT(k) (++sst[k])
#define STATG(k) atomicAdd(a.counters + (k), 1ull)   // where the tasks of the literal kernel come from: counters[9..14] (unused slots of k_seed's own)
#else
#define STAT(k) ((void)0)
#define STATG(k) ((void)0)
#endif
#ifndef CS_SPEC
#define CS_SPEC 1          // 0: never take the speculative unique-match path (literal sweeps only)
#endif

template <int RW>
__device__ __forceinline__ void seed_body(const DevIndex &I, const SeedArgs &a)
{
	extern __shared__ uint4 s_list[];                     // [CS_LIST_SMEM][CS_SEED_BLOCK] interval lists
	uint64_t *s_rd = reinterpret_cast<uint64_t*>(s_list + CS_LIST_SMEM * CS_SEED_BLOCK);       // [RW][CS_SEED_BLOCK] packed read
	uint32_t *s_nm = reinterpret_cast<uint32_t*>(s_rd + (RW ? RW : 1) * CS_SEED_BLOCK);        // [RW][CS_SEED_BLOCK] N mask
	const int t = threadIdx.x;
	const size_t nthreads = (size_t)gridDim.x * CS_SEED_BLOCK;
	const size_t gtid = (size_t)blockIdx.x * CS_SEED_BLOCK + t;
	cs_mem_t *const my = a.thread_mems + gtid * a.mem_cap;
	const cs_seed_opt_t opt = a.opt;
	// the filter is usable only if a filtered match is certain to be shorter than min_seed_len
	const int prune_k = (I.pt_k > 0 && opt.min_seed_len >= (int)I.pt_k) ? (int)I.pt_k : 0;
	const bool utext = I.text != nullptr;                 // unique-match fast paths available (2-bit text + sampled inverse SA)
	// (in call mode k_seed_fast has already tried the speculative route on every call it hands over)
	const bool can_spec = CS_SPEC && utext && prune_k > 0 && (int)I.kt_depth >= 2 && (int)I.kt_depth < prune_k && a.defer_q == nullptr;

	uint32_t n_ext = 0, n_call = 0, n_two = 0, n_probe = 0, ext_mark = 0;
	uint32_t n_req = 0;                                   // executed gathers besides Occ sectors and filter words (table, SA, inverse SA, text)
#ifdef CS_STATS
	uint32_t sst[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
	int st = ST_FETCH;
	uint32_t rd = 0, cur_q = 0; int len = 0;
	const uint64_t *pw = nullptr;                         // this read, 2-bit packed, in global memory
	uint32_t nmem = 0, old_n = 0, r2k = 0;
	int round = 1;
	int x = 0, i = 0, bi = 0, ret = 0;
	uint64_t c0 = 0, c1 = 0, c2 = 0; uint32_t cend = 0;   // current interval (forward ik / backward p)
	int n = 0, lo = 0, j = 0, w = 0;
	bool pushed = false; uint64_t last_sz = 0;
	uint64_t min_intv = 1;
	uint32_t call_nmem = 0; int last_start = 0;
	uint32_t kmask = 0xffffffffu;                         // bit d-1: a forward match of d < K bases may be pushed
	uint64_t tpos = 0;                                    // text position: of q[i] (ST_TXT_*), of q[bi+1] (ST_BTX_*, ST_ROW_* after them)
	int c = 0;
	bool need = false, jumped = false, spec = false;
	int rowm = 0;                                         // ST_ROW_*: 1 = x[0] wanted, 2 = x[1] wanted, 4 = forward pass (push), else emit
	bool err_list = false, err_mem = false;

	auto list_put = [&](int idx, uint4 v) {
		if (idx < CS_LIST_SMEM) s_list[idx * CS_SEED_BLOCK + t] = v;
		else if ((uint32_t)(idx - CS_LIST_SMEM) < a.spill_cap) a.spill[(size_t)(idx - CS_LIST_SMEM) * nthreads + gtid] = v;
		else err_list = true;
	};
	auto list_get = [&](int idx) -> uint4 {
		if (idx < CS_LIST_SMEM) return s_list[idx * CS_SEED_BLOCK + t];
		if ((uint32_t)(idx - CS_LIST_SMEM) < a.spill_cap) return a.spill[(size_t)(idx - CS_LIST_SMEM) * nthreads + gtid];
		return make_uint4(0, 0, 0, 0);
	};
	auto emit = [&](uint64_t x0, uint64_t x1, uint64_t x2, uint32_t start, uint32_t end) {
		if (nmem < a.mem_cap) {
			uint4 *p = reinterpret_cast<uint4*>(my + nmem);
			p[0] = make_uint4((uint32_t)x0, (uint32_t)(x0 >> 32), (uint32_t)x1, (uint32_t)(x1 >> 32));
			p[1] = make_uint4((uint32_t)x2, (uint32_t)(x2 >> 32), end, start);
		} else err_mem = true;
		++nmem;
	};
	// the read in flight: 2 bits per base + N mask
	auto rd_word = [&](int tt, const uint64_t *gp, uint32_t wi) -> uint64_t {
		if (RW) return s_rd[wi * CS_SEED_BLOCK + tt];
		return __ldg(gp + wi);
	};
	auto nm_word = [&](int tt, const uint64_t *gp, uint32_t wi) -> uint32_t {
		if (RW) return s_nm[wi * CS_SEED_BLOCK + tt];
		return __ldg(a.nmask + (gp - a.packed) + wi);
	};
	auto base_at = [&](int pos) -> int { // nt4 code of q[pos]: 0..3, or 4 for an ambiguous base
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		return ((nm_word(t, pw, wi) >> sh) & 1) ? 4 : (int)((rd_word(t, pw, wi) >> (2 * sh)) & 3);
	};
	auto key_of = [&](int tt, const uint64_t *gp, int pos, int cnt) -> uint64_t { // cnt (< 32) bases from pos, base j at bits 2j
		uint32_t wi = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31) * 2;
		uint64_t v = rd_word(tt, gp, wi) >> sh;
		if (sh) v |= rd_word(tt, gp, wi + 1) << (64 - sh);
		return v & ((1ull << (2 * cnt)) - 1);
	};
	auto nmask_window = [&](int pos) -> uint32_t { // bit j: q[pos + j] is ambiguous or past the end of the read
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		uint32_t m = nm_word(t, pw, wi) >> sh;
		if (sh) m |= nm_word(t, pw, wi + 1) << (32 - sh);
		return m;
	};
	auto read_window = [&](int pos) -> uint64_t { // the 32 bases from pos, base j at bits 2j
		uint32_t wi = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31) * 2;
		uint64_t v = rd_word(t, pw, wi) >> sh;
		if (sh) v |= rd_word(t, pw, wi + 1) << (64 - sh);
		return v;
	};
	auto has_n = [&](int tt, const uint64_t *gp, int pos, int cnt) -> bool { // any N / out-of-read base in [pos, pos+cnt), cnt < 32
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		uint32_t m = nm_word(tt, gp, wi) >> sh;
		if (sh) m |= nm_word(tt, gp, wi + 1) << (32 - sh);
		return (m & ((1u << cnt) - 1u)) != 0;
	};
	auto set_intv = [&]() { // bwt_set_intv, bwt.h:82, for q[x]
		int b = base_at(x);
		c0 = l2_at(I, b) + 1; c1 = l2_at(I, 3 - b) + 1; c2 = l2_at(I, b + 1) - l2_at(I, b);
		i = x + 1; n = 0; jumped = false;
	};
	auto start_call = [&](int pivot, uint64_t mi) { // bwt_smem1a prologue, bwt.c:295-302
		x = pivot; min_intv = mi < 1 ? 1 : mi;
		set_intv();
		call_nmem = 0; ret = x + 1; ext_mark = n_ext;
		spec = can_spec && min_intv == 1;
		STAT(4); if (spec) STAT(5);
		st = prune_k ? ST_FILTER : ST_FWD;
	};
	// kv_push(curr, ik) of the forward pass (bwt.c:312,317,321) -- unless the filter proved the match useless
	auto fwd_push = [&](int end) {
		ret = end;                                       // bwt.c:323: the longest forward match ends here
		int d = end - x;
		if (d >= prune_k || ((kmask >> (d - 1)) & 1)) list_put(n++, pack_entry(c0, c1, c2, (uint32_t)end));
	};
	// a match [bi+1, cend) ends at this sweep: it is an SMEM only if no longer match survived the
	// sweep and it is not contained in the previous one (bwt.c:332-336)
	auto mem_candidate = [&]() {
		if (call_nmem == 0 || bi + 1 < last_start) {
			++call_nmem; last_start = bi + 1;
			if ((int)cend - (bi + 1) >= opt.min_seed_len) emit(c0, c1, c2, (uint32_t)(bi + 1), cend);
		}
	};
	// Speculative mode (see "unique-match paths" below): the forward pass pushed nothing.  If the literal
	// pass would not have pushed anything either, the call is over; otherwise redo it literally.
	// Forward jump (result-neutral): no match shorter than J = (first depth the filter lets through) can be pushed,
	// so the table entry of q[x, x+J) is fetched directly instead of one depth per trip.  In speculative mode
	// nothing is pushed before the match is unique, so J is the table depth.  If that string occurs fewer
	// than min_intv times the call restarts one base at a time (see `jumped`).
	auto try_jump = [&]() {
		int J = (spec || kmask == 0) ? 32 : __ffs((int)kmask);
		if (J > (int)I.kt_depth) J = (int)I.kt_depth;
		if (J >= 2 && J < prune_k && !has_n(t, pw, x, J)) { i = x + J - 1; jumped = true; }
	};
	auto spec_abort = [&]() { spec = false; n_ext = ext_mark; set_intv(); try_jump(); st = ST_FWD; need = false; };
	auto spec_stop = [&](int end) { // the forward match [x, end) ended before it became unique, or is shorter than K
		int d = end - x;
		if (d < prune_k && (kmask & ((1u << d) - 1u)) == 0) { ret = end; st = ST_CALL_DONE; need = false; STAT(7); }   // nothing pushable: bwt.c:322 with curr->n == 0
		else { spec_abort(); STAT(8); }
	};
	// end of the forward pass at read position `end` (N, read end): bwt.c:317,321
	auto fwd_end = [&](int end) {
		if (spec) spec_stop(end);
		else { fwd_push(end); st = ST_BWD_INIT; need = false; }
	};
	// row of the suffix starting at text position p: nearest sample to the right, then `steps` LF steps back
	auto isa_near = [&](uint64_t p, uint64_t &row, int &steps) {
		const uint64_t smask = (1ull << I.isa_shift) - 1;
		uint64_t jj = (p + smask) & ~smask;
		if (jj > I.seq_len) jj = I.seq_len;
		row = jj == I.seq_len ? 0ull : gather_u64(I.isa + (jj >> I.isa_shift));   // the '$' suffix is row 0
		steps = (int)(jj - p); ++n_req;
	};

	for (;;) {
		// ---- divergent bookkeeping: advance this lane's state machine until it needs the warp.  With
		//      CS_BK_QUORUM > 1 the warp enters this section only when that many lanes wait for it (or
		//      nobody has convergent work), so its instructions are issued for several lanes at once ----
		bool bk = !need && st != ST_IDLE && st != ST_FILTER;
		if (CS_BK_QUORUM > 1) {
			const unsigned waiting = __ballot_sync(0xffffffffu, bk);
			const unsigned busy = __ballot_sync(0xffffffffu, need || st == ST_FILTER);
			if (busy && __popc(waiting) < CS_BK_QUORUM) bk = false;
		}
		if (bk) STAT(12);
		while (bk && !need && st != ST_IDLE && st != ST_FILTER) {
			switch (st) {
			case ST_FETCH: {
				rd = atomicAdd(a.next_read, 1u);
				uint4 item = make_uint4(0, 0, 0, 0);
				if (a.defer_q) { // call mode: only the calls k_seed_fast handed over
					if (rd >= *a.n_lit) { st = ST_IDLE; break; }
					cur_q = a.lit_q[rd]; item = a.defer_q[cur_q]; rd = item.x;
				} else if (rd >= a.n_reads) { st = ST_IDLE; break; }
				uint32_t o = a.off[rd];
				len = (int)(a.off[rd + 1] - o);
				pw = a.packed + ((uint64_t)((o + a.off_bias) >> 5) + 2ull * rd);
				if (RW) {
					const uint32_t *gn = a.nmask + (pw - a.packed);
					const uint32_t nw = ((uint32_t)len >> 5) + 2;
#pragma unroll
					for (uint32_t wi = 0; wi < (RW ? RW : 1); ++wi) {
						s_rd[wi * CS_SEED_BLOCK + t] = wi < nw ? __ldg(pw + wi) : 0ull;
						s_nm[wi * CS_SEED_BLOCK + t] = wi < nw ? __ldg(gn + wi) : 0xffffffffu;
					}
				}
				nmem = 0; round = 1; x = 0; err_list = err_mem = false;
				st = ST_R1_PIVOT;
				if (a.defer_q) { // one call: a first-pass pivot (then the second pass of what it finds) or a second-pass pivot
					round = (int)((item.y >> 16) & 3); old_n = 0; r2k = 0;
					start_call((int)(item.y & 0xffff), (uint64_t)item.z);
				}
			} break;
			case ST_R1_PIVOT: // first pass of mem_collect_intv, bwamem.c:226-236
				while (x < len && base_at(x) > 3) ++x;
				if (x >= len) { old_n = nmem; r2k = 0; st = ST_R2_NEXT; }
				else start_call(x, 1);
				break;
			case ST_FWD: { // forward extension, bwt.c:304-321
				int b = i < len ? base_at(i) : 4;
				if (b > 3) fwd_end(i);
				else { c = 3 - b; need = true; }
			} break;
			case ST_BWD_INIT: // bwt.c:322-326; list[n-1] is the longest match that was kept
				if (n == 0) st = ST_CALL_DONE;
				else { bi = x - 1; lo = 0; st = ST_BWD_SWEEP; }
				break;
			case ST_BWD_SWEEP: { // one value of i in bwt.c:326
				c = bi < 0 ? -1 : base_at(bi);
				if (c > 3) c = -1;
				if (c < 0) { // every interval ends here; only the longest can be a new SMEM (bwt.c:331-337)
					unpack_entry(list_get(n - 1), c0, c1, c2, cend);
					mem_candidate();
					st = ST_CALL_DONE;
				} else {
					j = n - 1; w = n; pushed = false; st = ST_BWD_ENTRY;
					if (utext && n - lo == 1 && min_intv == 1) { // one interval left: load it here, maybe take the text path
						unpack_entry(list_get(j), c0, c1, c2, cend);
						if (c2 == 1) st = ST_BTX_SA;
						need = true;
					}
				}
			} break;
			case ST_BWD_ENTRY:
				if (j < lo) {
					if (!pushed) st = ST_CALL_DONE;
					else { lo = w; --bi; st = ST_BWD_SWEEP; }
				} else { unpack_entry(list_get(j), c0, c1, c2, cend); need = true; }
				break;
			case ST_CALL_DONE:
				if (round != 1) st = ST_R2_NEXT;
				else if (a.defer_q) { old_n = nmem; r2k = 0; st = ST_R2_NEXT; }   // call mode: k_seed_fast goes on with the first pass
				else { x = ret; st = ST_R1_PIVOT; }
				break;
			case ST_R2_NEXT: { // second pass, bwamem.c:238-249 (the third pass runs in k_seed_r3)
				st = ST_READ_DONE;
				uint32_t lim = old_n < a.mem_cap ? old_n : a.mem_cap;
				while (r2k < lim) {
					const uint4 *p = reinterpret_cast<const uint4*>(my + r2k);
					uint4 v = p[1];
					++r2k;
					int s = (int)v.w, e = (int)v.z;
					uint64_t sz = (uint64_t)v.x | ((uint64_t)v.y << 32);
					if (e - s < opt.split_len || sz > (uint64_t)opt.split_width) continue;
					round = 2;
					start_call((s + e) >> 1, sz + 1);
					break;
				}
			} break;
			case ST_READ_DONE: {
				uint32_t cnt = nmem;
				if (err_mem || err_list) { cnt = 0; atomicMin(a.error, CS_E_READ_OVERFLOW); }
				unsigned long long o = atomicAdd(a.pool_used, (unsigned long long)cnt);
				if (o + cnt > a.pool_cap) { cnt = 0; atomicMin(a.error, CS_E_OVERFLOW); }
				if (a.defer_q) { a.x_off[cur_q] = o; a.x_n[cur_q] = cnt; }
				else { a.read_pool_off[rd] = o; a.read_n_mems[rd] = cnt; }
				const uint4 *src = reinterpret_cast<const uint4*>(my);
				uint4 *dst = reinterpret_cast<uint4*>(a.pool + o);
				for (uint32_t m = 0; m < 2 * cnt; ++m) dst[m] = src[m];
				st = ST_FETCH;
			} break;
			}
		}

		// ---- explicit reconvergence: all 32 lanes meet here every trip; nobody leaves early ----
		if (__all_sync(0xffffffffu, st == ST_IDLE)) break;
		__syncwarp();   // vote intrinsics are not memory barriers: make every lane's read words visible to the warp
		if ((t & 31) == 0) STAT(0);

		// ---- occurrence filter, served by the whole warp for one requesting lane at a time: lane l
		//      tests the window that ends l+1 bases after the pivot ----
		for (unsigned req = __ballot_sync(0xffffffffu, st == ST_FILTER); req; req &= req - 1) {
			const int owner = __ffs(req) - 1, lane = t & 31;
			const int ox = __shfl_sync(0xffffffffu, x, owner);
			const uint64_t *opw = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)pw, owner));
			const uint32_t omin = (uint32_t)__shfl_sync(0xffffffffu, (unsigned)(min_intv > 3 ? 4 : min_intv), owner);
			const int ot = t - lane + owner;
			const int ws = ox + 1 + lane - prune_k;           // window [ws, ws + K) ends lane+1 bases after the pivot
			bool keep = false;
			if (lane + 1 < prune_k && ws >= 0 && !has_n(ot, opw, ws, ox - ws)) { // else it ends at the read start / an N first
				const uint64_t key = key_of(ot, opw, ws, prune_k);
				const uint32_t cnt = (gather_u32(I.pt + (key >> 4)) >> (2 * ((uint32_t)key & 15))) & 3;
				keep = cnt == 3 || cnt >= omin;              // 3 == "3 or more"
				++n_probe;
			}
			const unsigned km = __ballot_sync(0xffffffffu, keep);
			if (lane == owner) {
				STAT(13);
				kmask = km; st = ST_FWD;
				try_jump();
			}
		}
		if (!need) continue;

		// ---- unique-match paths (result-neutral).  Once an interval holds ONE occurrence (x[2] == 1, min_intv ==
		//      1) every further bwt_extend only asks "does the next read base equal the next text base", so the
		//      extension is done 32 bases per step against the 2-bit text at the occurrence (one SA gather), and the
		//      coordinate that moves (x[1] forward, x[0] backward) is rebuilt at the end from the sampled inverse SA
		//      plus < 2^isa_shift LF steps.
		//      Forward (ST_TXT_*): x[0] stays the row of the occurrence, nothing is pushed while the size stays 1
		//      (bwt.c:311), the pass ends at the first mismatch / N / end.
		//      Backward (ST_BTX_*), one interval left in the sweep (bwt.c:329-341 with prev->n == 1): x[1], x[2] stay,
		//      the sweep that fails (read start, N, mismatch, text start) makes the SMEM.
		//      Speculative call (`spec`): the forward pass pushes nothing until the match is unique, and the backward
		//      pass follows only that longest match L = [x, cend) to the position f where it fails.  Every shorter
		//      match the literal pass would have pushed contains L's occurrence, so it survives every sweep L
		//      survives and cannot become an SMEM before f (curr->n > 0, bwt.c:332); at f it is rejected as
		//      contained (bwt.c:333) unless it survives f.  All of them start with q[f, f+K) once extended to f
		//      (their length is >= dlow, the first depth the filter lets through, and x + dlow - f >= K is checked),
		//      so if that K-mer does not occur in the text (one filter probe) none survives: the call's only
		//      SMEM is [f+1, cend).  Otherwise the call is redone literally (spec_abort). ----
		if (st >= ST_TXT_SA && st <= ST_ROW_LF) {
			bool fin = false;
			STAT(3); if (st == ST_TXT_CMP) STAT(15); if (st == ST_ROW_LF) STAT(11);
			if (st == ST_TXT_SA) {
				tpos = gather_u64(I.sa + c0) + (uint64_t)(i - x); ++n_req;
				j = 0; st = ST_TXT_CMP;
			} else if (st == ST_TXT_CMP) {
				n_req += 1u + ((tpos & 31) != 0);
				const uint64_t diff = read_window(i) ^ packed_window(I.text, tpos);
				const uint32_t nmw = nmask_window(i);
				uint32_t m = diff ? (uint32_t)(__ffsll((long long)diff) - 1) >> 1 : 32u;
				const uint32_t nn = nmw ? (uint32_t)__ffs((int)nmw) - 1u : 32u;
				const uint64_t left = I.seq_len - tpos;
				if (nn < m) m = nn;
				if (left < m) m = (uint32_t)left;
				i += (int)m; tpos += m; j += (int)m;
				if (m < 32) { // the forward pass ends here: at an N, the read end, the text end, or a mismatch (child size 0 != 1)
					n_ext += (unsigned)j + ((i < len && base_at(i) <= 3) ? 1u : 0u);   // bwt.c:306-320: no bwt_extend at an N / the read end
					if (!spec) {
						if (j > 0) { rowm = 2 | 4; st = ST_ROW_ISA; }
						else { fwd_push(i); st = ST_BWD_INIT; need = false; }
					} else if (i - x < prune_k && !((kmask >> (i - x - 1)) & 1)) spec_stop(i);   // L itself would not be pushed
					else { // L = [x, i): follow it backward from x - 1
						rowm = j > 0 ? 2 : 0;
						cend = (uint32_t)i; ret = i; bi = x - 1; tpos -= (uint64_t)(i - x); j = 0; st = ST_BTX_CMP;
					}
				}
			} else if (st == ST_BTX_SA) {
				tpos = gather_u64(I.sa + c0); ++n_req;
				j = 0; rowm = 0; st = ST_BTX_CMP;
			} else if (st == ST_BTX_CMP) {
				uint32_t cnt = 32, m = 0;
				if ((uint32_t)(bi + 1) < cnt) cnt = (uint32_t)(bi + 1);
				if (tpos < cnt) cnt = (uint32_t)tpos;
				if (cnt) { // the cnt bases q[bi-cnt+1 .. bi] against the cnt text bases before tpos, compared from the top
					const int sr = bi + 1 - (int)cnt;
					n_req += 1u + (((tpos - cnt) & 31) != 0);
					const uint64_t diff = (read_window(sr) ^ packed_window(I.text, tpos - cnt)) << (2 * (32 - cnt));
					const uint32_t nmw = nmask_window(sr) << (32 - cnt);
					m = diff ? (uint32_t)__clzll((long long)diff) >> 1 : 32u;
					const uint32_t nn = nmw ? (uint32_t)__clz((int)nmw) : 32u;
					if (nn < m) m = nn;
					if (cnt < m) m = cnt;
				}
				bi -= (int)m; tpos -= m; j += (int)m;
				if (m < 32) { // bi is the first position that does not extend the match
					const bool ext_fails = bi >= 0 && base_at(bi) <= 3;   // bwt.c:330: no bwt_extend at the read start / an N
					n_ext += (unsigned)j + (ext_fails ? 1u : 0u);
					bool ok = true;
					if (spec && ext_fails) { // could a shorter match survive position bi?
						const int dlow = kmask ? __ffs((int)kmask) : prune_k;
						ok = false; STAT(10);
						if (x + dlow - bi >= prune_k && !has_n(t, pw, bi, prune_k)) {
							const uint64_t key = key_of(t, pw, bi, prune_k);
							ok = ((gather_u32(I.pt + (key >> 4)) >> (2 * ((uint32_t)key & 15))) & 3) == 0;
							++n_probe;
#ifdef CS_STATS
							--sst[10]; if (!ok) STAT(9);
#endif
						}
					}
					if (!ok) spec_abort();
					else {
						if (spec) STAT(6);
						st = ST_CALL_DONE; need = false;
						if (call_nmem == 0 || bi + 1 < last_start) { // bwt.c:332-336
							++call_nmem; last_start = bi + 1;
							if ((int)cend - (bi + 1) >= opt.min_seed_len) {
								if (j > 0) rowm |= 1;
								if (rowm) { st = ST_ROW_ISA; need = true; }
								else emit(c0, c1, c2, (uint32_t)(bi + 1), cend);
							}
						}
					}
				}
			} else if (st == ST_ROW_ISA) {
				int w0 = 0, w1 = 0;
				if (rowm & 1) isa_near(tpos, c0, w0);                                       // row of the suffix at the match start
				if (rowm & 2) isa_near(I.seq_len - tpos - ((rowm & 4) ? 0ull : (uint64_t)((int)cend - bi - 1)), c1, w1);   // ... of its reverse complement
				w = w0 | (w1 << 8);
				if (w == 0) fin = true; else st = ST_ROW_LF;
			} else { // ST_ROW_LF: one LF step (row of the preceding suffix) on each coordinate still walking
				if (w & 0xff) { c0 = dev_lf(I, c0); w -= 1; ++n_req; }
				if (w >> 8) { c1 = dev_lf(I, c1); w -= 256; ++n_req; }
				if (w == 0) fin = true;
			}
			if (fin) {
				if (rowm & 4) { fwd_push(i); st = ST_BWD_INIT; }
				else { emit(c0, c1, c2, (uint32_t)(bi + 1), cend); st = ST_CALL_DONE; }
				need = false;
			}
			continue;
		}

		// ---- the one convergent, memory-bound step: bwt_extend of (c0,c1,c2) by base c ----
		const int is_back = (st == ST_BWD_ENTRY);
		// look-ahead base for the step after this one
		const int pf_idx = is_back ? bi - 1 : i + 1;
		uint32_t nb = 4;
		if (pf_idx >= 0 && pf_idx < len) nb = (uint32_t)base_at(pf_idx);
		uint64_t o0, o1, o2;
		{
			// the extended string is q[x..i] (forward) or q[bi..cend) (backward); short strings come from
			// the top-of-search table (one 16-byte gather, no Occ sectors), the rest from the FM-index
			const int s_beg = is_back ? bi : x;
			const int new_len = is_back ? (int)cend - bi : i + 1 - x;
			++n_ext;
			if (new_len <= (int)I.kt_depth) { kt_lookup(I, (uint32_t)new_len, key_of(t, pw, s_beg, new_len), o0, o1, o2); ++n_req; STAT(1); }
			else { uint32_t two; dev_extend(I, c0, c1, c2, c, is_back, o0, o1, o2, two); ++n_call; n_two += two; STAT(2); }
		}

		if (!is_back) { // ST_FWD, bwt.c:311-315
			if (jumped) { // this was the table entry of q[x, i]: J = i + 1 - x bases fetched at once
				jumped = false;
				if (o2 < min_intv) { i = x + 1; --n_ext; need = false; STAT(14); continue; }   // too rare: redo this call one base at a time
				n_ext += (unsigned)(i - x - 1); c2 = o2;
			}
			if (o2 != c2) {
				if (!spec) fwd_push(i);
				if (o2 < min_intv) { if (spec) spec_stop(i); else st = ST_BWD_INIT; need = false; }
			}
			if (need) {
				c0 = o0; c1 = o1; c2 = o2; ++i;
				if (i >= len || nb > 3) fwd_end(i);
				else if (utext && c2 == 1 && min_intv == 1) st = ST_TXT_SA;   // unique from here on
				else c = 3 - (int)nb;
			}
		} else { // ST_BWD_ENTRY, bwt.c:331-341
			if (o2 < min_intv) {
				if (!pushed) mem_candidate();
			} else if (!pushed || o2 != last_sz) {
				list_put(--w, pack_entry(o0, o1, o2, cend));
				pushed = true; last_sz = o2;
			}
			--j;
			if (j >= lo) unpack_entry(list_get(j), c0, c1, c2, cend);       // next interval of this sweep
			else if (pushed && bi >= 1 && nb < 4) {                         // next sweep, one base further left
				lo = w; --bi; c = (int)nb;
				j = n - 1; w = n; pushed = false;
				unpack_entry(list_get(j), c0, c1, c2, cend);
				if (utext && n - lo == 1 && c2 == 1 && min_intv == 1) st = ST_BTX_SA;   // one occurrence left: compare against the text
			} else need = false;                                            // ST_BWD_ENTRY finishes the sweep on the slow path
		}
	}
	if (n_ext) atomicAdd(a.counters + 0, (unsigned long long)n_ext);
	if (n_call) atomicAdd(a.counters + 1, (unsigned long long)n_call);
	if (n_two) atomicAdd(a.counters + 2, (unsigned long long)n_two);
	if (n_probe) atomicAdd(a.counters + 3, (unsigned long long)n_probe);
	n_req += n_call + n_two + n_probe;
	if (n_req) atomicAdd(a.req + 2, (unsigned long long)n_req);
#ifdef CS_STATS
	for (int k = 0; k < 16; ++k) if (sst[k]) atomicAdd(a.counters + 4 + k, (unsigned long long)sst[k]);
#endif
}

__global__ void __launch_bounds__(CS_SEED_BLOCK, CS_SEED_MINBLOCKS) k_seed(DevIndex I, SeedArgs a) { seed_body<CS_READ_SMEM>(I, a); }
// reads longer than 32 * CS_READ_SMEM bases: the packed read stays in global memory
__global__ void __launch_bounds__(CS_SEED_BLOCK, CS_SEED_MINBLOCKS) k_seed_long(DevIndex I, SeedArgs a) { seed_body<0>(I, a); }

// Warp-aggregated work distribution: the lanes of `want` get consecutive values of *ctr with ONE atomic per
// warp (the counters of a batch share a sector; one atomic per lane serialises in a single L2 slice).
// Must be called by all 32 lanes.
__device__ __forceinline__ uint32_t warp_take(uint32_t *ctr, bool want)
{
	const unsigned m = __ballot_sync(0xffffffffu, want);
	if (!m) return 0;
	const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(m));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
// ... and `cnt` consecutive slots each (cnt = 0 for lanes that want none): exclusive warp scan + one atomic
__device__ __forceinline__ unsigned long long warp_alloc(unsigned long long *ctr, uint32_t cnt)
{
	const int lane = threadIdx.x & 31;
#ifdef CS_EMUL   // one-lane warps (tests/emul/seed_emul.cpp)
	return cnt ? atomicAdd(ctr, (unsigned long long)cnt) : 0ull;
#endif
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
	const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0;
	if (total == 0) return 0;
	if (lane == 31) base = atomicAdd(ctr, (unsigned long long)total);
	base = __shfl_sync(0xffffffffu, base, 31);
	return base + (incl - cnt);
}

// ---------------------------------------------------------------------------------------------
// Fast seeding kernel (passes 1 and 2 for reads whose every bwt_smem1a call is "simple").
//
// k_seed above is a trip-synchronous state machine: every warp trip pays for every divergent path
// some lane is on, which makes it issue-bound (profiles/r01_ncu_k_seed_final_*).  This kernel is
// CALL-synchronous instead: every lane runs one whole bwt_smem1a call per iteration of the outer
// loop, and all lanes walk through the same phases (filter probes -> table jump -> FM extends until
// the match is unique -> text comparison forward and backward -> one K-mer probe -> rows from the
// sampled inverse SA).  It never builds an interval list.  It relies on the argument spelled out at
// "unique-match paths" in seed_body: the call's only SMEM is the backward extension of the longest
// forward match L when L is one occurrence and the K-mer at the position where L fails does not
// occur in the text; and a call whose forward pass could not push anything returns nothing.  The
// second-pass call of a one-occurrence SMEM is answered where the SMEM is found, from the repeat
// lengths at its text position (DevIndex::rep), without the FM-index or the filter.  Any CALL that
// does not fit (a shorter match could survive, pass 2 needs a list, scratch overflow) is queued as
// (read, pivot, min_intv, pass): for k_seed_walk when its list can be walked entry by entry (see
// there), else for k_seed, which executes it literally -- and, for a first-pass call, the
// second-pass calls of the SMEMs it finds (bwamem.c:238-249 depend only on their own SMEM).  The first pass goes on here: the next pivot is the end of the longest forward
// match (bwt.c:323), which this kernel knows exactly.  The collect pass merges a read's chain of
// deferred results with the ones written here.  Results are bit-identical by construction; only the
// share of deferred calls depends on the data (about 0.6 per read on the i.i.d. 3.1 Gbp
// configuration -- a random 19-mer has a second occurrence there with probability 2 % --, most
// calls on repeat-rich references).
// ---------------------------------------------------------------------------------------------
template <int RW>
__device__ __forceinline__ void seed_fast_body(const DevIndex &I, const SeedArgs &a)
{
	extern __shared__ uint4 s_dyn[];
	uint64_t *s_rd = reinterpret_cast<uint64_t*>(s_dyn);                                   // [RW][CS_FAST_BLOCK] packed read
	uint32_t *s_nm = reinterpret_cast<uint32_t*>(s_rd + RW * CS_FAST_BLOCK);              // [RW][CS_FAST_BLOCK] N mask
	const int t = threadIdx.x;
	const size_t gtid = (size_t)blockIdx.x * CS_FAST_BLOCK + t;
	cs_mem_t *const my = a.thread_mems + gtid * a.mem_cap;
	const cs_seed_opt_t opt = a.opt;
	const int K = (int)I.pt_k, kd = (int)I.kt_depth;       // host guarantees 2 <= kd < K <= min_seed_len, K <= 19

	uint32_t n_ext = 0, n_call = 0, n_two = 0, n_probe = 0;
	uint32_t r_ext = 0, r_call = 0;                        // of the call in flight: counted only if it is not deferred
	uint32_t n_req = 0;                                    // EXECUTED gathers other than filter words (Occ sectors, table, SA, inverse SA, text), deferred calls included
#ifdef CS_STATS
	uint32_t sst[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
	bool have = false, exhausted = false;
	uint32_t last_q = 0xffffffffu;                        // most recent deferred call of the read in flight (chain head)
	uint32_t rd = 0; int len = 0;
	int round = 1, x = 0;
	uint32_t nmem = 0, old_n = 0, r2k = 0;
	long long diag = 0; bool have_diag = false; int junk_at = -1;   // diagonal of the last one-occurrence SMEM this lane stored for the read; where its forward match failed on a base
	uint64_t p2done = 0;                                  // bit m: the second-pass call of my[m] has been dealt with where the SMEM was found (repeat lengths)

	auto rd_word = [&](uint32_t wi) -> uint64_t { return s_rd[wi * CS_FAST_BLOCK + t]; };
	auto nm_word = [&](uint32_t wi) -> uint32_t { return s_nm[wi * CS_FAST_BLOCK + t]; };
	auto base_at = [&](int pos) -> int {
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		return ((nm_word(wi) >> sh) & 1) ? 4 : (int)((rd_word(wi) >> (2 * sh)) & 3);
	};
	auto read_window = [&](int pos) -> uint64_t { // the 32 bases from pos, base j at bits 2j
		uint32_t wi = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31) * 2;
		uint64_t v = rd_word(wi) >> sh;
		if (sh) v |= rd_word(wi + 1) << (64 - sh);
		return v;
	};
	auto nmask_window = [&](int pos) -> uint32_t { // bit j: q[pos + j] is ambiguous or past the end of the read
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		uint32_t m = nm_word(wi) >> sh;
		if (sh) m |= nm_word(wi + 1) << (32 - sh);
		return m;
	};
	auto key_of = [&](int pos, int cnt) -> uint64_t { return read_window(pos) & ((1ull << (2 * cnt)) - 1); };   // cnt < 32
	auto has_n = [&](int pos, int cnt) -> bool { return (nmask_window(pos) & ((1u << cnt) - 1u)) != 0; };       // cnt < 32
	auto pt_count = [&](uint64_t key) -> uint32_t { return (gather_u32(I.pt + (key >> 4)) >> (2 * ((uint32_t)key & 15))) & 3; };
	auto isa_near = [&](uint64_t p, uint64_t &row, int &steps) {
		const uint64_t smask = (1ull << I.isa_shift) - 1;
		uint64_t jj = (p + smask) & ~smask;
		if (jj > I.seq_len) jj = I.seq_len;
		row = jj == I.seq_len ? 0ull : gather_u64(I.isa + (jj >> I.isa_shift));
		steps = (int)(jj - p); ++n_req;
	};
	// hand the call (pivot, min_intv) of the read in flight to the literal kernel; the queue slot is taken at the top
	// of the next iteration, for the whole warp at once
	uint32_t pend_y = 0, pend_z = 0, pend_bits = 0;           // pend_y != 0: a call waits to be queued
	auto defer_call = [&](int pivot, uint64_t mi) {
		pend_y = (uint32_t)pivot | ((uint32_t)round << 16); pend_z = (uint32_t)mi;
	};
	// ... or, when its list can only hold the few short matches `bits` (depth of the longest forward match: d), to k_seed_walk
	auto defer_walk = [&](int pivot, uint64_t mi, int d, uint32_t bits) {
		pend_y = (uint32_t)pivot | ((uint32_t)round << 16) | ((uint32_t)d << 18) | 0x80000000u; pend_z = (uint32_t)mi; pend_bits = bits;
	};
	// ... or, when the longest forward match L has K or more bases (so it is pushed) but is not one occurrence: L itself, with its
	// interval (x2 == 0: not known, k_seed_walk extends to it), walked first; the K-mer probe at the position where it fails then
	// decides about every other entry of the list (bit 27; d in bits 18-22 and 24-26)
	// (the interval waits in this thread's slot of thread_lx, not in registers)
	auto defer_walk_l = [&](int pivot, int pass, uint64_t mi, int d, uint64_t x0, uint64_t x1, uint64_t x2) {
		pend_y = (uint32_t)pivot | ((uint32_t)pass << 16) | (((uint32_t)d & 31u) << 18) | (1u << 23) | (((uint32_t)d >> 5) << 24) | (1u << 27) | 0x80000000u;
		pend_z = (uint32_t)mi; pend_bits = 0; a.thread_lx[gtid] = pack_entry(x0, x1, x2, 0);
	};

	for (;;) {
		// ---- queue the calls deferred in the previous iteration (one atomic per warp) ----
		{
			const uint32_t q = warp_take(a.n_defer, pend_y != 0);
			if (pend_y != 0) {
				// (past the capacity nothing is stored: the host sees n_defer > defer_cap and reruns the batch through k_seed alone)
				if (q < a.defer_cap) {
					a.defer_q[q] = make_uint4(rd, pend_y, pend_z, last_q);
					if (pend_y >> 31) a.defer_bits[q] = pend_bits;
					if ((pend_y >> 27) & 1) a.defer_lx[q] = a.thread_lx[gtid];
					last_q = q;
				}
			}
			// the calls for the literal kernel are also listed on their own, so that it does not scan the walk tasks
			const bool lit = pend_y != 0 && !(pend_y >> 31) && q < a.defer_cap;
			const uint32_t li = warp_take(a.n_lit, lit);
			if (lit) a.lit_q[li] = q;
			pend_y = 0;
		}
		// ---- pick this lane's next call: (cx, cmin).  Reads are handed out, and finished reads get their place in
		//      the pool, for the whole warp at once. ----
		bool active = false; int cx = 0; uint64_t cmin = 1;
		for (;;) {
			const bool want = !exhausted && !have;
			const uint32_t take = warp_take(a.next_read + 2, want);
			if (want) {
				rd = take;
				if (rd >= a.n_reads) exhausted = true;
				else {
					const uint32_t o = a.off[rd];
					len = (int)(a.off[rd + 1] - o);
					const uint64_t w0 = (uint64_t)((o + a.off_bias) >> 5) + 2ull * rd;
					const uint32_t nw = ((uint32_t)len >> 5) + 2;
#pragma unroll
					for (uint32_t wi = 0; wi < RW; ++wi) {
						s_rd[wi * CS_FAST_BLOCK + t] = wi < nw ? __ldg(a.packed + w0 + wi) : 0ull;
						s_nm[wi * CS_FAST_BLOCK + t] = wi < nw ? __ldg(a.nmask + w0 + wi) : 0xffffffffu;
					}
					nmem = 0; round = 1; x = 0; have = true; last_q = 0xffffffffu; p2done = 0; have_diag = false; junk_at = -1;
				}
			}
			bool finished = false;
			if (have && !active) {
				if (round == 1) { // first pass of mem_collect_intv, bwamem.c:226-236
					while (x < len && base_at(x) > 3) ++x;
					if (x < len) { cx = x; cmin = 1; active = true; }
					else { old_n = nmem; r2k = 0; round = 2; }
				}
				if (!active) {
					while (r2k < old_n) { // second pass, bwamem.c:238-249
						const uint32_t k2 = r2k++;
						if (k2 < 64 && ((p2done >> k2) & 1)) continue;           // answered when the SMEM was found
						const uint4 v = reinterpret_cast<const uint4*>(my + k2)[1];
						const int s = (int)v.w, e = (int)v.z;
						const uint64_t sz = (uint64_t)v.x | ((uint64_t)v.y << 32);
						if (e - s < opt.split_len || sz > (uint64_t)opt.split_width) continue;
						cx = (s + e) >> 1; cmin = sz + 1; active = true;
						break;
					}
					finished = !active;
				}
			}
			{ // the finished reads
				uint32_t cnt = finished ? nmem : 0;
				const unsigned long long o = warp_alloc(a.pool_used, cnt);
				if (finished) {
					if (o + cnt > a.pool_cap) { cnt = 0; atomicMin(a.error, CS_E_OVERFLOW); }
					a.read_pool_off[rd] = o; a.read_n_mems[rd] = cnt;
					const uint4 *src = reinterpret_cast<const uint4*>(my);
					uint4 *dst = reinterpret_cast<uint4*>(a.pool + o);
					for (uint32_t m = 0; m < 2 * cnt; ++m) dst[m] = src[m];
					a.read_last_q[rd] = last_q;
					have = false;
				}
			}
			if (!__any_sync(0xffffffffu, finished && !exhausted)) break;   // nobody is waiting for another read
		}
		if (__all_sync(0xffffffffu, !active)) break;
		if ((t & 31) == 0) STAT(0);
		if (!active) continue;     // exhausted lanes wait at the vote above
		STAT(1); if (cmin != 1) STAT(2);
		r_ext = r_call = 0;

		// ---- occurrence filter (as in seed_body): bit e-1 <=> the K-mer ending e bases after the pivot occurs
		//      >= min(cmin, 3) times, i.e. a forward match of e bases may be pushed.  Evaluated lazily: only the
		//      windows a decision below depends on are probed (these probes are the largest share of this
		//      kernel's DRAM traffic).  All probes of one request are independent loads. ----
		int e0;                                                     // first window that starts inside the N-free run left of the pivot
		{
			const int s0 = cx >= 31 ? cx - 31 : 0, cl = cx - s0;     // that run, capped at 31 bases
			int nl = 0;
			if (cl > 0) {
				const uint32_t m = (nmask_window(s0) & ((1u << cl) - 1u)) << (32 - cl);
				nl = m ? __clz((int)m) : cl;
			}
			e0 = K - nl < 1 ? 1 : K - nl;
		}
		auto probe_bits = [&](int elo, int ehi) -> uint32_t {       // filter bits of the windows elo..ehi
			const uint32_t omin = cmin > 3 ? 4u : (uint32_t)cmin;
			uint32_t cnt[18], mask = 0;
			if (elo < e0) elo = e0;
#pragma unroll
			for (int u = 0; u < 18; ++u) {
				const int e = u + 1;
				cnt[u] = 0;
				if (e >= elo && e <= ehi && e < K) { cnt[u] = pt_count(key_of(cx + e - K, K)); ++n_probe; }
			}
#pragma unroll
			for (int u = 0; u < 18; ++u) if (cnt[u] == 3 || (cnt[u] != 0 && cnt[u] >= omin)) mask |= 1u << u;
			return mask;
		};

		// ---- forward pass (bwt.c:304-321) without pushes: until the match is one occurrence, dies, or hits an N / the end ----
		uint64_t c0 = 0, c1 = 0, c2 = 1;
		int i = cx + 1, udepth = 0;                                 // udepth: bases after which one occurrence was left
		bool unique = false, spec = false, fwd_go = true;
		uint64_t tp0 = 0, bw0 = 0, tpos = 0; int jf = 0; uint32_t cnt0 = 0;
		// First-pass calls, speculatively: a read with substitutions continues on the DIAGONAL (text position minus read position)
		// of its last one-occurrence match.  If q[cx, ..) equals the text there for more than R = rep[that position] bases, the
		// match is one occurrence from R + 1 bases on -- at that very position -- and everything the FM-index would have been
		// asked (table jump, extends down to one row, the SA gather) is known: three independent gathers (repeat length, text
		// window, the window before it for the backward pass) instead of a chain of five to eight.  The rows of the SMEM come
		// from the inverse SA at the end, as for every match followed through the text.  Not tried at the position where the
		// last match failed on a base (q differs from the text there by construction).
		if (CS_SPEC_DIAG && I.rep && cmin == 1 && have_diag && cx != junk_at) {
			const long long pp = diag + cx;
			if (pp >= 0 && (unsigned long long)pp < I.seq_len) {
				const uint64_t p1 = (uint64_t)pp;
				const uint32_t R = gather_u8(I.rep + p1);
				const uint64_t win = packed_window(I.text, p1);
				cnt0 = 32; if ((uint32_t)cx < cnt0) cnt0 = (uint32_t)cx; if (p1 < cnt0) cnt0 = (uint32_t)p1;
				if (cnt0) bw0 = packed_window(I.text, p1 - cnt0);
				n_req += 2u + ((p1 & 31) != 0) + (cnt0 ? 1u + (((p1 - cnt0) & 31) != 0) : 0u);
				const uint64_t diff = read_window(cx) ^ win;
				const uint32_t nmw = nmask_window(cx);
				uint32_t m = diff ? (uint32_t)(__ffsll((long long)diff) - 1) >> 1 : 32u;
				const uint32_t nn = nmw ? (uint32_t)__ffs((int)nmw) - 1u : 32u;
				const uint64_t lft = I.seq_len - p1;
				if (nn < m) m = nn;
				if (lft < m) m = (uint32_t)lft;
				if (R < 255u && m >= R + 1u) {
					STAT(7);
					spec = true; unique = true; udepth = (int)R + 1; tp0 = p1;
					i = cx + (int)m; jf = (int)m; tpos = p1 + m; fwd_go = m == 32u;
				} else STAT(9);
			}
		}
		if (!spec) {
			{
				const int b = base_at(cx);
				c0 = l2_at(I, b) + 1; c1 = l2_at(I, 3 - b) + 1; c2 = l2_at(I, b + 1) - l2_at(I, b);   // bwt_set_intv, bwt.h:82
			}
			if (!has_n(cx, kd)) { // q[cx, cx+kd) is inside the read and unambiguous: its table entry directly
				uint64_t o0, o1, o2;
				kt_lookup(I, (uint32_t)kd, key_of(cx, kd), o0, o1, o2); ++n_req;
				if (o2 >= cmin) { c0 = o0; c1 = o1; c2 = o2; i = cx + kd; r_ext += (uint32_t)(kd - 1); }
			}
			for (;;) {
				if (c2 == 1 && cmin == 1) { unique = true; udepth = i - cx; break; }
				const int b = i < len ? base_at(i) : 4;
				if (b > 3) break;
				const int new_len = i + 1 - cx;
				uint64_t o0, o1, o2;
				++r_ext;
				if (new_len <= kd) { kt_lookup(I, (uint32_t)new_len, key_of(cx, new_len), o0, o1, o2); ++n_req; }
				else { uint32_t two; dev_extend(I, c0, c1, c2, 3 - b, 0, o0, o1, o2, two); ++r_call; n_two += two; n_req += 1u + two; }
				if (o2 < cmin) break;                                   // bwt.c:313: this extension fails, the match ends at i
				c0 = o0; c1 = o1; c2 = o2; ++i;
			}
			if (unique) {
				tp0 = gather_u64(I.sa + c0); ++n_req;                        // text position of q[cx]
				// the text before the occurrence is wanted by the backward pass below: fetch its first window together with the forward one
				cnt0 = 32; if ((uint32_t)cx < cnt0) cnt0 = (uint32_t)cx; if (tp0 < cnt0) cnt0 = (uint32_t)tp0;
				if (cnt0) { bw0 = packed_window(I.text, tp0 - cnt0); n_req += 1u + (((tp0 - cnt0) & 31) != 0); }
				tpos = tp0 + (uint64_t)(i - cx);
			}
		}
		// unique from here on: compare against the text at the occurrence, 32 bases per step
		int fail_at = -1;                                           // where the text comparison failed on a base of the read
		if (unique) {
			// CS_FWD_WIN windows of 32 bases are fetched per trip (all their loads in flight together).  More than one looked
			// attractive (at 1 % substitutions a unique match runs on for ~100 bases, and the trip count differs from lane to
			// lane) but measured slower: see CS_FWD_WIN in cs_kernels.cuh.
			for (bool go = fwd_go; go; ) {
				const uint64_t left = I.seq_len - tpos;
				uint64_t room = (uint64_t)(len - i); if (left < room) room = left;            // bases that can still match
				const uint32_t nwin = room >= 32u * CS_FWD_WIN ? (uint32_t)CS_FWD_WIN : (uint32_t)((room + 31) >> 5);
				const uint64_t w = tpos >> 5; const uint32_t sh = ((uint32_t)tpos & 31) * 2;
				uint64_t tw[CS_FWD_WIN + 1];
#pragma unroll
				for (int k = 0; k <= CS_FWD_WIN; ++k) tw[k] = ((uint32_t)k < nwin || ((uint32_t)k == nwin && sh && nwin)) ? gather_u64(I.text + w + k) : 0ull;
				n_req += nwin + ((sh && nwin) ? 1u : 0u);
				if (nwin == 0) break;                                                         // read end / text end: nothing left to compare
#pragma unroll
				for (int k = 0; k < CS_FWD_WIN; ++k) {
					if (!go) break;
					if ((uint32_t)k >= nwin) { go = false; break; }
					const uint64_t win = sh ? (tw[k] >> sh) | (tw[k + 1] << (64 - sh)) : tw[k];
					const uint64_t diff = read_window(i) ^ win;
					const uint32_t nmw = nmask_window(i);
					uint32_t m = diff ? (uint32_t)(__ffsll((long long)diff) - 1) >> 1 : 32u;
					const uint32_t nn = nmw ? (uint32_t)__ffs((int)nmw) - 1u : 32u;
					const uint64_t lft = I.seq_len - tpos;
					if (nn < m) m = nn;
					if (lft < m) m = (uint32_t)lft;
					i += (int)m; tpos += m; jf += (int)m;
					if (m < 32) go = false;
				}
			}
			const bool on_base = i < len && base_at(i) <= 3;            // the match failed on a base of the read (not at an N / the read end)
			r_ext += (uint32_t)jf + (on_base ? 1u : 0u) - (spec ? 1u : 0u);   // (speculative: the comparison started at the pivot itself)
			if (on_base && tpos < I.seq_len) fail_at = i;
		}
		const int end = i, d = end - cx;                            // the longest forward match is L = [cx, end)
		if (round == 1) x = end;                                    // next pivot (bwt.c:323, bwamem.c:228)

		// ---- what would the literal pass have pushed?  Matches of >= K bases always; shorter ones by the filter. ----
		uint32_t kmask = 0;
		if (d < K) {
			kmask = probe_bits(1, d);
			if (kmask == 0) { STAT(3); n_ext += r_ext; n_call += r_call; continue; }   // nothing: the call returns no SMEM
		}
		if (cmin != 1 || !(d >= K || ((kmask >> (d - 1)) & 1))) { // pass 2 with a list, or L itself not pushed
			if (cmin != 1) STAT(4); else STAT(5);
			if (d < K && cmin < 0x80000000ull) defer_walk(cx, cmin, d, kmask);   // (however many entries: k_seed_walk's K-mer probe after the longest one leaves few)
			else if (CS_WALK_L && d >= K && d < 256 && cmin < 0x80000000ull) defer_walk_l(cx, round, cmin, d, c0, c1, c2);
			else { defer_call(cx, cmin); STATG(9); }
			continue;
		}
		if (!unique) { // a short L that still has several occurrences (the next mismatch came before the match was unique): its
			STAT(6);    // backward sweeps are bwt_extend steps like those of any other short match -- k_seed_walk does them
			if (d < K) { defer_walk(cx, cmin, d, kmask); pend_y |= 1u << 23; }   // bit 23: L first, then the K-mer probe decides about the rest
			else if (CS_WALK_L && d < 256) defer_walk_l(cx, round, cmin, d, c0, c1, c2);
			else { defer_call(cx, cmin); STATG(10); }
			continue;
		}

		// ---- backward: follow L alone (one occurrence, text position tp0) to the position where it fails ----
		int bi = cx - 1, jb = 0;
		uint64_t tb = tp0;                                          // text position of q[bi+1]
		for (bool first_w = true; ; first_w = false) {
			uint32_t cnt = 32, m = 0;
			if ((uint32_t)(bi + 1) < cnt) cnt = (uint32_t)(bi + 1);
			if (tb < cnt) cnt = (uint32_t)tb;
			if (cnt) {
				const int sr = bi + 1 - (int)cnt;
				const uint64_t tw = first_w ? bw0 : packed_window(I.text, tb - cnt);   // (first window: cnt == cnt0)
				if (!first_w) n_req += 1u + (((tb - cnt) & 31) != 0);
				const uint64_t diff = (read_window(sr) ^ tw) << (2 * (32 - cnt));
				const uint32_t nmw = nmask_window(sr) << (32 - cnt);
				m = diff ? (uint32_t)__clzll((long long)diff) >> 1 : 32u;
				const uint32_t nn = nmw ? (uint32_t)__clz((int)nmw) : 32u;
				if (nn < m) m = nn;
				if (cnt < m) m = cnt;
			}
			bi -= (int)m; tb -= m; jb += (int)m;
			if (m < 32) break;
		}
		const bool ext_fails = bi >= 0 && base_at(bi) <= 3;         // bwt.c:330: no bwt_extend at the read start / an N
		r_ext += (uint32_t)jb + (ext_fails ? 1u : 0u);
		uint32_t rest = 0;                                          // pushed matches that may survive position bi: k_seed_walk walks them one by one
		if (ext_fails) { // Those of >= K - (cx - bi) bases all start with q[bi, bi+K) there: if it does not occur they end there, contained in L
			const int need_d = K - (cx - bi);
			if (need_d > 1) rest = (d < K ? kmask : probe_bits(1, need_d - 1)) & ((1u << (need_d - 1)) - 1u);
			bool absent = false;
			if (!has_n(bi, K)) { absent = pt_count(key_of(bi, K)) == 0; ++n_probe; }
			if (!absent) { // every pushed match has to be looked at.  Beyond K bases only size changes before the match was unique are pushed:
				STAT(8);    // none if it was unique by then (otherwise the literal kernel takes the call)
				if (udepth > K) { defer_call(cx, cmin); STATG(10); continue; }
				rest = d < K ? kmask & ~(1u << (d - 1)) : probe_bits(1, K - 1);
				if (udepth >= 1 && udepth < 32) rest &= (1u << (udepth - 1)) - 1u;   // from udepth on the size stays 1: not pushed (bwt.c:311)
			}
		}
		STAT(10);
		if (end - (bi + 1) >= opt.min_seed_len && nmem >= a.mem_cap) { defer_call(cx, cmin); STATG(11); continue; }   // scratch full: the literal kernel stores it
		// (each entry is a walk of its own in one lane of k_seed_walk: calls with many of them would hold their warp up)
		if (__popc(rest) > CS_WALK_MAX) { defer_call(cx, cmin); STATG(11); continue; }
		// they come after L in the sweep order: k_seed_walk starts its containment test from L's start
		if (rest) defer_walk(cx, cmin, d < 31 ? d : 31, rest | ((uint32_t)(bi + 2) << 18));
		n_ext += r_ext; n_call += r_call;
		if (end - (bi + 1) < opt.min_seed_len) continue;            // bwamem.c:231-233,247
		STAT(11);
		diag = (long long)tb - (bi + 1); have_diag = true; junk_at = fail_at;   // hints for the speculation above: nothing else depends on them
		// ---- the second-pass call of this SMEM (bwamem.c:238-249: pivot in its middle, min_intv 2), answered here from the repeat
		//      lengths when it can be: the SMEM has ONE occurrence, at text position tb, so q == T there.  The call's forward
		//      match is R = rep[pivot's text position] bases long if that ends inside the SMEM (its interval holds >= 2 rows up
		//      to R bases, one row after); the K-mer windows that decide which shorter matches the literal pass would push
		//      (probe_bits above) lie inside the SMEM too when the pivot is >= K-1 bases from its start, and occur twice iff their
		//      own rep is >= K.  Neither the FM-index nor the filter is read.  The loads are issued now and used after the
		//      inverse-SA lookups below. ----
		bool p2_try = false; int cx2 = 0; uint32_t ro = 0;
		uint64_t rw0 = 0, rw1 = 0, rw2 = 0, rw3 = 0;
		if (I.rep && round == 1 && nmem < 64 && end - (bi + 1) >= opt.split_len && opt.split_width >= 1) {
			cx2 = (bi + 1 + end) >> 1;
			if (cx2 - (bi + 1) >= K - 1) {
				const uint64_t lo = tb + (uint64_t)(cx2 - (bi + 1)) - (uint64_t)(K - 1);   // text position of the first window, q[cx2+1-K, cx2+1)
				const uint64_t *rp = reinterpret_cast<const uint64_t*>(I.rep) + (lo >> 3);
				ro = (uint32_t)lo & 7;                                                       // (K + 7 <= 32 bytes: four words always hold them)
				rw0 = gather_u64(rp); rw1 = gather_u64(rp + 1); rw2 = gather_u64(rp + 2); rw3 = gather_u64(rp + 3);
				n_req += 1u + ((((uint32_t)lo & 31) + (uint32_t)K > 32u) ? 1u : 0u);        // sectors
				p2_try = true;
			}
		}
		// ---- the coordinates that moved by text comparison: x[0] backward, x[1] forward ----
		{
			int w0 = 0, w1 = 0;
			if (jb > 0 || spec) isa_near(tb, c0, w0);
			if (jf > 0) isa_near(I.seq_len - tb - (uint64_t)(end - bi - 1), c1, w1);
			while (w0 | w1) {
				if (w0) { c0 = dev_lf(I, c0); --w0; ++n_req; }
				if (w1) { c1 = dev_lf(I, c1); --w1; ++n_req; }
			}
		}
		if (p2_try) {
			const uint32_t Kq = (uint32_t)K * 0x01010101u;
			auto ge_k = [&](uint64_t w) -> uint32_t {   // bit i: byte i of w >= K
				const uint32_t lo = (__vcmpgeu4((uint32_t)w, Kq) >> 7) & 0x01010101u, hi = (__vcmpgeu4((uint32_t)(w >> 32), Kq) >> 7) & 0x01010101u;
				return (((lo * 0x01020408u) >> 24) & 0xfu) | ((((hi * 0x01020408u) >> 24) & 0xfu) << 4);
			};
			const uint32_t flags = (ge_k(rw0) | (ge_k(rw1) << 8) | (ge_k(rw2) << 16) | (ge_k(rw3) << 24)) >> ro;   // bit j: the window of e = j + 1 occurs twice
			const uint32_t ri = ro + (uint32_t)K - 1u;
			const uint64_t rw = ri < 8 ? rw0 : ri < 16 ? rw1 : ri < 24 ? rw2 : rw3;
			const int R = (int)((rw >> (8 * (ri & 7))) & 0xff);
			if (R >= 1 && R < 255 && R < end - cx2) { // the forward match of the call is q[cx2, cx2+R), and it fails on a base of the read
				const uint32_t km = R < K ? flags & ((1u << R) - 1u) : 0u;
				bool done = true;
				if (R < K && km == 0) { STAT(12); n_ext += (uint32_t)R; }              // nothing pushable: the call returns no SMEM
				else if (pend_y != 0) done = false;                                     // (one call can wait to be queued: the regular path takes this one)
				else if (R < K) { STAT(13); pend_y = (uint32_t)cx2 | (2u << 16) | ((uint32_t)R << 18) | 0x80000000u; pend_z = 2u; pend_bits = km; }
				else if (CS_WALK_L && R >= K) { STAT(14); defer_walk_l(cx2, 2, 2, R, 0, 0, 0); }   // (its interval is not known here)
				else { STAT(14); STATG(14); pend_y = (uint32_t)cx2 | (2u << 16); pend_z = 2u; }
				if (done) p2done |= 1ull << nmem;
			} else STAT(15);
		}
		{
			uint4 *p = reinterpret_cast<uint4*>(my + nmem);
			p[0] = make_uint4((uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32));
			p[1] = make_uint4((uint32_t)c2, (uint32_t)(c2 >> 32), (uint32_t)end, (uint32_t)(bi + 1));
			++nmem;
		}
	}
	if (n_ext) atomicAdd(a.counters + 0, (unsigned long long)n_ext);
	if (n_call) atomicAdd(a.counters + 1, (unsigned long long)n_call);
	if (n_two) atomicAdd(a.counters + 2, (unsigned long long)n_two);
	if (n_probe) atomicAdd(a.counters + 3, (unsigned long long)n_probe);
	n_req += n_probe;
	if (n_req) atomicAdd(a.req + 0, (unsigned long long)n_req);
#ifdef CS_STATS
	for (int k = 0; k < 16; ++k) if (sst[k]) atomicAdd(a.counters + 20 + k, (unsigned long long)sst[k]);
#endif
}

__global__ void __launch_bounds__(CS_FAST_BLOCK, CS_FAST_MINBLOCKS) k_seed_fast(DevIndex I, SeedArgs a) { seed_fast_body<CS_READ_SMEM>(I, a); }

// ---------------------------------------------------------------------------------------------
// Walk kernel: the deferred calls whose interval list can be walked ENTRY BY ENTRY.
//
// In bwt_smem1a's backward phase a list entry's fate does not depend on the others except through
// the start of the last SMEM found: a longer entry is a right-extension of a shorter one, so it dies
// no later; hence when an entry fails to extend no longer entry is alive (curr->n == 0, bwt.c:332), and
// duplicates dropped by the size test (bwt.c:338) would fail at the same base as the entry that
// shadows them and be rejected as contained (bwt.c:333).  So each entry is walked backward alone,
// longest first, with plain bwt_extend steps (table or FM-index), and the containment test is applied
// in that order.  An entry is in the list only if the forward pass pushed it: the size changed at
// the next base (bwt.c:311-312) or it is the forward pass's last interval (bwt.c:317,321).
//
// What k_seed_fast hands over, with the filter bits of the entries E_e = q[cx, cx+e), e < K (one bit per
// K-mer window that occurs often enough, §"Occurrence filter" at k_seed):
//   - calls whose longest forward match L has fewer than K bases (L not pushable itself, L not yet one
//     occurrence, or a second-pass call);
//   - calls whose L has K or more bases but several occurrences: L comes first, with its interval
//     (defer_lx; computed here when the call was answered from the repeat lengths), the other entries
//     are looked for only if the probe below leaves any;
//   - the shallow entries left over after k_seed_fast resolved L itself (ls0 != 0).
// After the LONGEST entry has been walked to the position bi where it fails, every other entry of
// >= K - (cx - bi) bases, extended that far, starts with q[bi, bi+K): if that K-mer occurs fewer than
// min_intv times they all end there, contained in the SMEM just found -- one filter probe instead of
// one walk per entry, which is what makes lists of any size acceptable (a call that starts in the
// middle of an error-free stretch has one entry per K-mer of the stretch).
// The second-pass calls of what a first-pass call finds here (bwamem.c:238-249) are run in place:
// forward pass, then the list.  One task per lane; all lanes run the same phases; the tasks come in
// descending order of their expected length (k_walk_*), so that a warp's 32 tasks are of one size.
// What does not fit -- entries that survive the probe (> CS_WALK_MAX of them), a walk of more than
// CS_WALK_STEPS bases, a full scratch -- goes to the literal kernel.
// ---------------------------------------------------------------------------------------------
#ifndef CS_WALK_STEPS
#define CS_WALK_STEPS 96
#endif
// The walk tasks of a batch in descending order of their expected length (counting sort over 64 classes: count, scan, scatter), so that
// the 32 tasks a warp of k_seed_walk runs in lock step are of one size: left in queue order, a warp waits for its longest task with
// 7 of 32 lanes busy (profiles/r02_ncu_all_kernels_base_*).  Expected length = extends of the longest entry (forward steps past the
// table, the pushed test, >= K - e backward steps for a window that occurs) + a few per further entry; 24 for a list that starts with L.
__device__ __forceinline__ uint32_t walk_cost_class(uint4 item, uint32_t bits, int K, int kd)
{
	uint32_t cost = ((item.y >> 27) & 1) ? 24u : 0u;
	const uint32_t nb = bits & 0x3ffffu;
	if (nb) {
		const int e = 32 - __clz((int)nb);
		cost += (uint32_t)((e > kd ? e - kd : 0) + 3 + (K > e ? K - e : 0)) + 4u * (uint32_t)(__popc(nb) - 1);
	}
	return cost < 63u ? cost : 63u;
}
__global__ void k_walk_count(DevIndex I, SeedArgs a, uint32_t *hist)
{
	__shared__ uint32_t s_h[64];
	if (threadIdx.x < 64) s_h[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t nq = *a.n_defer_fast < a.defer_cap ? *a.n_defer_fast : a.defer_cap;
	for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
		const uint4 item = a.defer_q[q];
		if (item.y >> 31) atomicAdd(&s_h[walk_cost_class(item, a.defer_bits[q], (int)I.pt_k, (int)I.kt_depth)], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 64 && s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_h[threadIdx.x]);
}
__global__ void k_walk_scan(uint32_t *hist, uint32_t *cursor, uint32_t *n_walk)
{ // one thread: 64 classes, the longest first
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		uint32_t acc = 0;
		for (int c = 63; c >= 0; --c) { cursor[c] = acc; acc += hist[c]; }
		*n_walk = acc;
	}
}
__global__ void k_walk_scatter(DevIndex I, SeedArgs a, uint32_t *cursor, uint32_t *order)
{
	__shared__ uint32_t s_h[64], s_base[64];
	const uint32_t nq = *a.n_defer_fast < a.defer_cap ? *a.n_defer_fast : a.defer_cap;
	const uint32_t tile = blockDim.x * 8;
	for (uint32_t t0 = blockIdx.x * tile; t0 < nq; t0 += gridDim.x * tile) { // a tile per CTA and trip: one global atomic per class and tile
		if (threadIdx.x < 64) s_h[threadIdx.x] = 0;
		__syncthreads();
		uint32_t cls[8], pos[8];
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const uint32_t q = t0 + k * blockDim.x + threadIdx.x;
			cls[k] = 0xffffffffu;
			if (q < nq) {
				const uint4 item = a.defer_q[q];
				if (item.y >> 31) { cls[k] = walk_cost_class(item, a.defer_bits[q], (int)I.pt_k, (int)I.kt_depth); pos[k] = atomicAdd(&s_h[cls[k]], 1u); }
			}
		}
		__syncthreads();
		if (threadIdx.x < 64) s_base[threadIdx.x] = s_h[threadIdx.x] ? atomicAdd(cursor + threadIdx.x, s_h[threadIdx.x]) : 0u;
		__syncthreads();
#pragma unroll
		for (int k = 0; k < 8; ++k) if (cls[k] != 0xffffffffu) order[s_base[cls[k]] + pos[k]] = t0 + k * blockDim.x + threadIdx.x;
		__syncthreads();
	}
}

#ifndef CS_R3_QUORUM
#define CS_R3_QUORUM 8
#endif
__global__ void __launch_bounds__(CS_FAST_BLOCK, CS_WALK_MINBLOCKS) k_seed_walk(DevIndex I, SeedArgs a)
{
	constexpr int RW = CS_READ_SMEM;
	extern __shared__ uint4 s_dyn[];
	uint64_t *s_rd = reinterpret_cast<uint64_t*>(s_dyn);
	uint32_t *s_nm = reinterpret_cast<uint32_t*>(s_rd + RW * CS_FAST_BLOCK);
	const int t = threadIdx.x;
	const size_t gtid = (size_t)blockIdx.x * CS_FAST_BLOCK + t;
	cs_mem_t *const my = a.thread_mems + gtid * a.mem_cap;
	const cs_seed_opt_t opt = a.opt;
	const int kd = (int)I.kt_depth, K = (int)I.pt_k;
	uint32_t n_ext = 0, n_call = 0, n_two = 0, n_req = 0;   // n_req: executed gathers (Occ sectors, table entries, filter words)
	bool exhausted = false;

	auto rd_word = [&](uint32_t wi) -> uint64_t { return s_rd[wi * CS_FAST_BLOCK + t]; };
	auto nm_word = [&](uint32_t wi) -> uint32_t { return s_nm[wi * CS_FAST_BLOCK + t]; };
	auto base_at = [&](int pos) -> int {
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		return ((nm_word(wi) >> sh) & 1) ? 4 : (int)((rd_word(wi) >> (2 * sh)) & 3);
	};
	auto key_of = [&](int pos, int cnt) -> uint64_t { // cnt < 32 bases from pos, base j at bits 2j
		uint32_t wi = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31) * 2;
		uint64_t v = rd_word(wi) >> sh;
		if (sh) v |= rd_word(wi + 1) << (64 - sh);
		return v & ((1ull << (2 * cnt)) - 1);
	};
	auto nmask_window = [&](int pos) -> uint32_t { // bit j: q[pos + j] is ambiguous or past the end of the read
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		uint32_t m = nm_word(wi) >> sh;
		if (sh) m |= nm_word(wi + 1) << (32 - sh);
		return m;
	};

	for (;;) {
		// ---- next walk task of the queue (tasks of the literal kernel are skipped) ----
		uint32_t q = 0; uint4 item = make_uint4(0, 0, 0, 0);
		bool active = false;
		for (;;) { // queue slots for the whole warp at once; lanes that drew a task of the literal kernel draw again
			const bool want = !exhausted && !active;
			const uint32_t take = warp_take(a.next_read + 3, want);
			if (want) { // the walk tasks k_seed_fast queued, longest first (k_walk_scatter); entries appended by this kernel are not among them
				if (take >= *a.n_walk) exhausted = true;
				else { q = a.walk_order[take]; item = a.defer_q[q]; active = true; }
			}
			if (!__any_sync(0xffffffffu, !exhausted && !active)) break;
		}
		if (__all_sync(0xffffffffu, !active)) break;
		const uint32_t rd = item.x;
		const int cx0 = (int)(item.y & 0xffff), round = (int)((item.y >> 16) & 3);
		const int d0 = (int)(((item.y >> 18) & 31) | (((item.y >> 24) & 7) << 5));
		const bool lfirst0 = active && ((item.y >> 27) & 1) != 0;   // the longest entry is L itself, K or more bases: not among the bits, its interval comes with the task
		const uint64_t cmin0 = item.z;
		uint32_t bits0 = active ? a.defer_bits[q] : 0u;
		const int ls00 = (int)(bits0 >> 18);                        // 1 + start of the SMEM k_seed_fast already found for this call (0: none)
		bits0 &= 0x3ffffu;
		uint64_t lx0 = 0, lx1 = 0, lx2 = 0;
		if (lfirst0) { uint32_t dummy; unpack_entry(a.defer_lx[q], lx0, lx1, lx2, dummy); }
		int len = 0;
		if (active) {
			const uint32_t o = a.off[rd];
			len = (int)(a.off[rd + 1] - o);
			const uint64_t w0 = (uint64_t)((o + a.off_bias) >> 5) + 2ull * rd;
			const uint32_t nw = ((uint32_t)len >> 5) + 2;
#pragma unroll
			for (uint32_t wi = 0; wi < RW; ++wi) {
				s_rd[wi * CS_FAST_BLOCK + t] = wi < nw ? __ldg(a.packed + w0 + wi) : 0ull;
				s_nm[wi * CS_FAST_BLOCK + t] = wi < nw ? __ldg(a.nmask + w0 + wi) : 0xffffffffu;
			}
		}
		uint32_t nm = 0, t_ext = 0, t_call = 0;
		// filter bits of the windows e = elo .. ehi of a call (pivot cx, min_intv cmin), as k_seed_fast's probe_bits: bit e-1 <=> the K-mer
		// ending e bases after the pivot occurs >= min(cmin, 3) times; windows that reach into an N / before the read start have no bit
		auto probe_bits_w = [&](int cx, uint64_t cmin, int elo, int ehi) -> uint32_t {
			int nl = 0;                                             // N-free run left of the pivot, capped at 31 bases
			{
				const int s0 = cx >= 31 ? cx - 31 : 0, cl = cx - s0;
				if (cl > 0) {
					const uint32_t m = (nmask_window(s0) & ((1u << cl) - 1u)) << (32 - cl);
					nl = m ? __clz((int)m) : cl;
				}
			}
			const int e0 = K - nl < 1 ? 1 : K - nl;
			const uint32_t omin = cmin > 3 ? 4u : (uint32_t)cmin;
			uint32_t mask = 0;
			if (elo < e0) elo = e0;
			if (ehi > K - 1) ehi = K - 1;
			for (int eb = 0; eb < 18; eb += 6) { // six independent gathers at a time
				uint32_t cnt[6];
#pragma unroll
				for (int u = 0; u < 6; ++u) {
					const int e = eb + u + 1;
					cnt[u] = 0;
					if (e >= elo && e <= ehi) { const uint64_t key = key_of(cx + e - K, K); cnt[u] = (gather_u32(I.pt + (key >> 4)) >> (2 * ((uint32_t)key & 15))) & 3; ++n_req; }
				}
#pragma unroll
				for (int u = 0; u < 6; ++u) if (cnt[u] == 3 || (cnt[u] != 0 && cnt[u] >= omin)) mask |= 1u << (eb + u);
			}
			return mask;
		};
		// ONE bwt_smem1a call whose forward pass is known (pivot cx, min_intv cmin, longest forward match of d bases): its list walked
		// entry by entry, longest first; SMEMs of >= min_seed_len bases are appended to my[nm..].  on: this lane has such a call.
		// Returns true if the call has to go to the literal kernel instead.
		// The loops are made warp-uniform with votes: left to themselves the lanes drift apart (their entries take the table or the
		// FM-index at different steps) and the hardware ends up running them one after the other.
		auto run_list = [&](bool on_call, int cx, uint64_t cmin, int d, uint32_t bits, bool lfirst, uint64_t l0, uint64_t l1, uint64_t l2, int ls0) -> bool {
			const bool lgiven = on_call && lfirst;                  // (the other entries of such a call are not known yet)
			bool first = ls0 == 0, punt = false;
			// After the LONGEST entry of a list has been walked to the position bi where it fails, every other entry of >= K - (cx - bi) bases,
			// extended that far, starts with q[bi, bi+K): one probe of that K-mer decides about all of them (below).  k_seed_fast has
			// already done that for the tasks that follow an SMEM of its own (ls0 != 0).
			bool lprobe = on_call && ls0 == 0;
			int last_start = ls0 - 1;
			if (!on_call) { bits = 0; lfirst = false; }
			while (__any_sync(0xffffffffu, (bits || lfirst) && !punt)) { // entries, longest first
				const bool on = (bits || lfirst) && !punt;
				int e = 0;
				uint64_t c0 = 0, c1 = 0, c2 = 0;
				bool go = false;
				if (on && lfirst) { // L = q[cx, cx+d), d >= K: always pushed (the forward pass's last interval, bwt.c:317,321)
					lfirst = false;
					e = d;
					c0 = l0; c1 = l1; c2 = l2;
					if (c2 == 0) { // interval not known (the call was answered from the repeat lengths): table, then forward bwt_extend steps
						kt_lookup(I, (uint32_t)kd, key_of(cx, kd), c0, c1, c2); ++n_req;
						for (int k = kd; k < e; ++k) {
							uint64_t o0, o1, o2; uint32_t two;
							dev_extend(I, c0, c1, c2, 3 - base_at(cx + k), 0, o0, o1, o2, two);
							c0 = o0; c1 = o1; c2 = o2; ++t_call; n_two += two; n_req += 1u + two;
						}
					}
					go = true;
				} else if (on) {
					e = 32 - __clz((int)bits);
					bits &= ~(1u << (e - 1));
					// the interval of q[cx, cx+e) ...
					kt_lookup(I, (uint32_t)(e < kd ? e : kd), key_of(cx, e < kd ? e : kd), c0, c1, c2); ++n_req;
					for (int k = kd; k < e; ++k) { // ... deeper than the table: forward bwt_extend steps
						uint64_t o0, o1, o2; uint32_t two;
						dev_extend(I, c0, c1, c2, 3 - base_at(cx + k), 0, o0, o1, o2, two);
						c0 = o0; c1 = o1; c2 = o2; ++t_call; n_two += two; n_req += 1u + two;
					}
					go = true;
					if (e < d) { // pushed only if the next forward step changed the size (bwt.c:311-312)
						uint64_t o0, o1, o2;
						if (e + 1 <= kd) { kt_lookup(I, (uint32_t)(e + 1), key_of(cx, e + 1), o0, o1, o2); ++n_req; }
						else { uint32_t two; dev_extend(I, c0, c1, c2, 3 - base_at(cx + e), 0, o0, o1, o2, two); ++t_call; n_two += two; n_req += 1u + two; }
						if (o2 == c2) go = false;
					}
				}
				const bool walked = go;
				// backward sweeps of this entry alone (bwt.c:326-345)
				int bi = cx - 1;
				for (int steps = 0; __any_sync(0xffffffffu, go); ++steps) {
					if (go) {
						const int b = bi >= 0 ? base_at(bi) : 4;
						if (b > 3) go = false;                              // read start / N: no bwt_extend (bwt.c:330)
						else if (steps >= CS_WALK_STEPS && ls0 == 0) { punt = true; go = false; }   // (a task that follows an SMEM of k_seed_fast is finished here)
						else {
							const int new_len = cx + e - bi;
							uint64_t o0, o1, o2;
							++t_ext;
							if (new_len <= kd) { kt_lookup(I, (uint32_t)new_len, key_of(bi, new_len), o0, o1, o2); ++n_req; }
							else { uint32_t two; dev_extend(I, c0, c1, c2, b, 1, o0, o1, o2, two); ++t_call; n_two += two; n_req += 1u + two; }
							if (o2 < cmin) go = false;                      // bwt.c:331
							else { c0 = o0; c1 = o1; c2 = o2; --bi; }
						}
					}
				}
				if (!walked || punt) continue;
				if (first || bi + 1 < last_start) { // bwt.c:332-336
					first = false; last_start = bi + 1;
					if (cx + e - (bi + 1) >= opt.min_seed_len) { // bwamem.c:231-233,247
						if (nm >= a.mem_cap) { punt = true; continue; }
						uint4 *p = reinterpret_cast<uint4*>(my + nm);
						p[0] = make_uint4((uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32));
						p[1] = make_uint4((uint32_t)c2, (uint32_t)(c2 >> 32), (uint32_t)(cx + e), (uint32_t)(bi + 1));
						++nm;
					}
				}
				if (lprobe) { // that was the longest entry: the K-mer probe at the position where it failed
					lprobe = false;
					if (bi < 0 || base_at(bi) > 3) bits = 0;            // read start / N: every interval ends here (bwt.c:331)
					else {
						const int need_d = K - (cx - bi);
						const uint32_t shallow = need_d > 1 ? bits & ((1u << (need_d - 1)) - 1u) : 0u;
						if (bits != shallow || lgiven) { // some entry is (may be) long enough for the probe to decide
							const uint64_t key = key_of(bi, K);
							// "does not occur" for a call with min_intv = cmin: fewer than cmin occurrences (the filter counts up to 3)
							const bool absent = !(nmask_window(bi) & ((1u << K) - 1u)) && ((gather_u32(I.pt + (key >> 4)) >> (2 * ((uint32_t)key & 15))) & 3) < (cmin > 3 ? 3u : (uint32_t)cmin);
							++n_req;
							if (absent) bits = shallow;                     // they all end at bi, contained in the SMEM just found; else every entry is walked
							else if (lgiven) punt = true;                   // ... which the literal kernel does: this call does not know them
						}
						if (lgiven && !punt && need_d > 1) bits = probe_bits_w(cx, cmin, 1, need_d - 1);   // the shallow entries of such a call
						if (__popc(bits) > CS_WALK_MAX) punt = true;        // too many for one lane: the literal kernel takes the call
					}
				}
			}
			return punt;
		};
		const bool punt = run_list(active, cx0, cmin0, d0, bits0, lfirst0, lx0, lx1, lx2, ls00);
		if (punt) { // the literal kernel takes it (it runs after this one)
			STATG(12);
			a.defer_q[q].y = item.y & 0x7fffffffu;
			a.lit_q[atomicAdd(a.n_lit, 1u)] = q;
		}
		const bool fin = active && !punt;
		// second-pass calls of what a first-pass call found (bwamem.c:238-249): forward pass here (table jump, then bwt_extend steps until
		// fewer than min_intv rows are left), then the list as above; the ones that do not fit go to the literal kernel as ordinary calls
		{
			const uint32_t nm_main = (fin && round == 1) ? nm : 0u;
			uint32_t fu = 0;
			bool fgo = nm_main > 0;
			while (__any_sync(0xffffffffu, fgo)) {
				int cx2 = 0; uint64_t cmin2 = 2; bool have = false;
				if (fgo) {
					while (fu < nm_main) {
						const uint4 v = reinterpret_cast<const uint4*>(my + fu)[1];
						++fu;
						const int s = (int)v.w, e = (int)v.z;
						const uint64_t sz = (uint64_t)v.x | ((uint64_t)v.y << 32);
						if (e - s < opt.split_len || sz > (uint64_t)opt.split_width) continue;
						cx2 = (s + e) >> 1; cmin2 = sz + 1; have = true;
						break;
					}
					if (!have) fgo = false;
				}
				uint64_t c0 = 0, c1 = 0, c2 = 0;
				int i = cx2 + 1;
				bool go = have;
				if (have) {
					STATG(13);
					const int b = base_at(cx2);
					c0 = l2_at(I, b) + 1; c1 = l2_at(I, 3 - b) + 1; c2 = l2_at(I, b + 1) - l2_at(I, b);   // bwt_set_intv, bwt.h:82
					if (!(nmask_window(cx2) & ((1u << kd) - 1u))) { // q[cx2, cx2+kd) is inside the read and unambiguous: its table entry directly
						uint64_t o0, o1, o2;
						kt_lookup(I, (uint32_t)kd, key_of(cx2, kd), o0, o1, o2); ++n_req;
						if (o2 >= cmin2) { c0 = o0; c1 = o1; c2 = o2; i = cx2 + kd; t_ext += (uint32_t)(kd - 1); }
					}
				}
				while (__any_sync(0xffffffffu, go)) { // bwt.c:304-321 without the pushes
					if (go) {
						const int b = i < len ? base_at(i) : 4;
						if (b > 3) go = false;
						else {
							const int new_len = i + 1 - cx2;
							uint64_t o0, o1, o2;
							++t_ext;
							if (new_len <= kd) { kt_lookup(I, (uint32_t)new_len, key_of(cx2, new_len), o0, o1, o2); ++n_req; }
							else { uint32_t two; dev_extend(I, c0, c1, c2, 3 - b, 0, o0, o1, o2, two); ++t_call; n_two += two; n_req += 1u + two; }
							if (o2 < cmin2) go = false;                     // bwt.c:313
							else { c0 = o0; c1 = o1; c2 = o2; ++i; }
						}
					}
				}
				const int d2 = i - cx2;
				uint32_t bits2 = 0;
				bool lf2 = false, lit2 = false;
				if (have) {
					if (d2 >= K) { if (d2 < 256) lf2 = true; else lit2 = true; }
					else bits2 = probe_bits_w(cx2, cmin2, 1, d2);
				}
				const uint32_t nm_before = nm;
				const bool punt2 = run_list(have && !lit2, cx2, cmin2, d2, bits2, lf2, c0, c1, c2, 0);
				if (have && (punt2 || lit2)) {
					nm = nm_before;
					const uint32_t q2 = atomicAdd(a.n_defer, 1u);
					if (q2 < a.defer_cap) {
						a.defer_q[q2] = make_uint4(rd, (uint32_t)cx2 | (2u << 16), (uint32_t)cmin2, atomicExch(a.read_last_q + rd, q2));
						a.lit_q[atomicAdd(a.n_lit, 1u)] = q2;
					}
				}
			}
		}
		if (fin) { n_ext += t_ext; n_call += t_call; }
		{
			uint32_t cnt = fin ? nm : 0;
			const unsigned long long o = warp_alloc(a.pool_used, cnt);
			if (fin) {
				if (o + cnt > a.pool_cap) { cnt = 0; atomicMin(a.error, CS_E_OVERFLOW); }
				a.x_off[q] = o; a.x_n[q] = cnt;
				const uint4 *src = reinterpret_cast<const uint4*>(my);
				uint4 *dst = reinterpret_cast<uint4*>(a.pool + o);
				for (uint32_t m = 0; m < 2 * cnt; ++m) dst[m] = src[m];
			}
		}
	}
	if (n_ext) atomicAdd(a.counters + 0, (unsigned long long)n_ext);
	if (n_call) atomicAdd(a.counters + 1, (unsigned long long)n_call);
	if (n_two) atomicAdd(a.counters + 2, (unsigned long long)n_two);
	if (n_req) atomicAdd(a.req + 1, (unsigned long long)n_req);
}

// ---------------------------------------------------------------------------------------------
// Third pass ("LAST-like", bwamem.c:253-268 + bwt_seed_strategy1, bwt.c:358-379) as its own kernel:
// forward-only chains with no interval lists, so it needs no shared memory and few registers and
// runs at full occupancy.  It is independent of passes 1-2; the collect pass merges and sorts.
// Read r may emit at most len/(min_seed_len+1) seeds; they go to r3_mems[off[r]/(k+1) + r ...].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) k_seed_r3(DevIndex I, SeedArgs a)
{
	const cs_seed_opt_t opt = a.opt;
	const uint32_t kp1 = (uint32_t)opt.min_seed_len + 1;
	unsigned long long n_ext = 0, n_call = 0, n_req = 0;
	bool idle = false, need = false;
	uint32_t rd = 0, nmem = 0; int len = 0, x = 0, i = 0, c = 0;
	const uint8_t *q = nullptr;
	const uint64_t *pw = nullptr; const uint32_t *pn = nullptr;
	cs_mem_t *out = nullptr;
	uint64_t c0 = 0, c1 = 0, c2 = 0;
	bool have_read = false;
	// a chain may start directly at depth `jump` from the top-of-search table: no seed can end before
	// min_seed_len + 1 bases (bwt.c:370), so the skipped intervals are never looked at
	const int jump = (int)I.kt_depth < opt.min_seed_len ? (int)I.kt_depth : opt.min_seed_len;

	for (;;) {
		while (!need && !idle) {
			if (!have_read) {
				rd = atomicAdd(a.next_read + 1, 1u);
				if (rd >= a.n_reads) { idle = true; break; }
				uint32_t o = a.off[rd];
				q = a.bases + o; len = (int)(a.off[rd + 1] - o);
				pw = a.packed + ((uint64_t)((o + a.off_bias) >> 5) + 2ull * rd); pn = a.nmask + ((uint64_t)((o + a.off_bias) >> 5) + 2ull * rd);
				out = a.r3_mems + ((uint64_t)(o / kp1) + rd);
				nmem = 0; x = 0; i = 0; have_read = true;
			}
			if (i <= x) { // need a new pivot
				while (x < len && q[x] > 3) ++x;
				if (x >= len) { a.r3_n_mems[rd] = nmem; have_read = false; continue; }
				if (jump >= 2 && !read_has_n(pn, x, jump)) { // q[x..x+jump) is inside the read and unambiguous
					kt_lookup(I, (uint32_t)jump, read_key(pw, x, jump), c0, c1, c2);
					i = x + jump; n_ext += (unsigned)(jump - 1); ++n_req;
				} else {
					int b = q[x];
					c0 = l2_at(I, b) + 1; c1 = l2_at(I, 3 - b) + 1; c2 = l2_at(I, b + 1) - l2_at(I, b);
					i = x + 1;
				}
			}
			// bwt.c:366-378
			if (i >= len) { a.r3_n_mems[rd] = nmem; have_read = false; }
			else if (q[i] > 3) { x = i + 1; i = 0; }
			else if (c2 == 0) { // children of an empty interval are empty: no memory access needed
				++n_ext;
				if (i - x >= opt.min_seed_len) { x = i + 1; i = 0; } else ++i;
			} else { c = 3 - q[i]; need = true; }
		}
		if (__all_sync(0xffffffffu, idle)) break;
		if (!need) continue;
		uint32_t nb = 4;
		if (i + 1 < len) nb = q[i + 1];
		uint64_t o0, o1, o2; uint32_t two;
		dev_extend(I, c0, c1, c2, c, 0, o0, o1, o2, two);
		++n_ext; ++n_call; n_req += 1u + two;
		if (o2 < (uint64_t)opt.max_mem_intv && i - x >= opt.min_seed_len) { // bwt.c:370-374
			if (o2 > 0) {
				uint4 *p = reinterpret_cast<uint4*>(out + nmem);
				p[0] = make_uint4((uint32_t)o0, (uint32_t)(o0 >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32));
				p[1] = make_uint4((uint32_t)o2, (uint32_t)(o2 >> 32), (uint32_t)(i + 1), (uint32_t)x);
				++nmem;
			}
			x = i + 1; i = 0; need = false;
		} else {
			c0 = o0; c1 = o1; c2 = o2; ++i;
			if (i >= len || nb > 3 || c2 == 0) need = false;
			else c = 3 - (int)nb;
		}
	}
	if (n_ext) atomicAdd(a.counters + 0, n_ext);
	if (n_call) atomicAdd(a.counters + 1, n_call);
	if (n_req) atomicAdd(a.req + 3, n_req);
}

// ---------------------------------------------------------------------------------------------
// Third pass, text-assisted (used together with k_seed_fast; k_seed_r3 above is the general kernel).
//
// bwt_seed_strategy1 (bwt.c:358-379) walks forward from x until the interval has fewer than
// max_mem_intv rows AND at least min_seed_len + 1 bases; on a mostly unique reference that is the
// (min_seed_len+1)-mer W = q[x, x+k+1) itself.  If W lies inside a first-pass SMEM with ONE
// occurrence (k_seed_fast left it in the read's pool entry, and its text position is SA[x[0]]), W
// occurs at a known text position P; and if the K-mer filter says that W's first K bases occur
// exactly once in the text, W occurs exactly once too: its bi-interval is (row of suffix P, row of
// the suffix where revcomp(W) starts, 1), two lookups in the sampled inverse SA plus a few LF steps
// instead of one table jump and k+1-13 bwt_extend calls with two Occ sectors each.  A chain that
// cannot reach k+1 bases (N, read end) or whose first K bases do not occur at all ends without a
// seed, with the same next x as the reference.  Every other chain (repeats,
// K > k, chains outside such an SMEM whose K-mer exists) is walked literally, as in k_seed_r3.
// One read per lane, one chain per iteration, all lanes in the same phase.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CS_FAST_BLOCK, CS_FAST_MINBLOCKS) k_seed_r3_fast(DevIndex I, SeedArgs a)
{
	constexpr int RW = CS_READ_SMEM;
	extern __shared__ uint4 s_dyn[];
	uint64_t *s_rd = reinterpret_cast<uint64_t*>(s_dyn);
	uint32_t *s_nm = reinterpret_cast<uint32_t*>(s_rd + RW * CS_FAST_BLOCK);
	const int t = threadIdx.x;
	const cs_seed_opt_t opt = a.opt;
	const int K = (int)I.pt_k, kd = (int)I.kt_depth, W = opt.min_seed_len + 1;   // host guarantees K <= min_seed_len
	const uint32_t kp1 = (uint32_t)W;
	const int jump = kd < opt.min_seed_len ? kd : opt.min_seed_len;
	uint32_t n_ext = 0, n_call = 0, n_probe = 0, n_req = 0;
	bool have = false, exhausted = false;
	bool parked = false;                                      // this lane's chain needs the literal walk and waits for company
	uint32_t rd = 0, nmem = 0, n12 = 0; int len = 0, x = 0;
	const cs_mem_t *pool12 = nullptr; cs_mem_t *out = nullptr;
	int ms = 0, me = 0; uint64_t mtb = 0;                     // the unique first-pass SMEM [ms, me) last used, at text position mtb

	auto rd_word = [&](uint32_t wi) -> uint64_t { return s_rd[wi * CS_FAST_BLOCK + t]; };
	auto nm_word = [&](uint32_t wi) -> uint32_t { return s_nm[wi * CS_FAST_BLOCK + t]; };
	auto base_at = [&](int pos) -> int {
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		return ((nm_word(wi) >> sh) & 1) ? 4 : (int)((rd_word(wi) >> (2 * sh)) & 3);
	};
	auto key_of = [&](int pos, int cnt) -> uint64_t {
		uint32_t wi = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31) * 2;
		uint64_t v = rd_word(wi) >> sh;
		if (sh) v |= rd_word(wi + 1) << (64 - sh);
		return v & ((1ull << (2 * cnt)) - 1);
	};
	auto nmask_window = [&](int pos) -> uint32_t {
		uint32_t wi = (uint32_t)pos >> 5, sh = (uint32_t)pos & 31;
		uint32_t m = nm_word(wi) >> sh;
		if (sh) m |= nm_word(wi + 1) << (32 - sh);
		return m;
	};
	auto isa_near = [&](uint64_t p, uint64_t &row, int &steps) {
		const uint64_t smask = (1ull << I.isa_shift) - 1;
		uint64_t jj = (p + smask) & ~smask;
		if (jj > I.seq_len) jj = I.seq_len;
		row = jj == I.seq_len ? 0ull : gather_u64(I.isa + (jj >> I.isa_shift));
		steps = (int)(jj - p); ++n_req;
	};
	auto put = [&](uint64_t x0, uint64_t x1, uint64_t x2, int start, int end) {
		uint4 *p = reinterpret_cast<uint4*>(out + nmem);
		p[0] = make_uint4((uint32_t)x0, (uint32_t)(x0 >> 32), (uint32_t)x1, (uint32_t)(x1 >> 32));
		p[1] = make_uint4((uint32_t)x2, (uint32_t)(x2 >> 32), (uint32_t)end, (uint32_t)start);
		++nmem;
	};

	for (;;) {
		// ---- this lane's next chain start x (bwamem.c:253-268) ----
		bool active = false;
		for (;;) { // reads are handed out for the whole warp at once
			const bool want = !exhausted && !have && !parked;
			const uint32_t take = warp_take(a.next_read + 1, want);
			if (want) {
				rd = take;
				if (rd >= a.n_reads) exhausted = true;
				else {
					const uint32_t o = a.off[rd];
					len = (int)(a.off[rd + 1] - o);
					const uint64_t w0 = (uint64_t)((o + a.off_bias) >> 5) + 2ull * rd;
					const uint32_t nw = ((uint32_t)len >> 5) + 2;
#pragma unroll
					for (uint32_t wi = 0; wi < RW; ++wi) {
						s_rd[wi * CS_FAST_BLOCK + t] = wi < nw ? __ldg(a.packed + w0 + wi) : 0ull;
						s_nm[wi * CS_FAST_BLOCK + t] = wi < nw ? __ldg(a.nmask + w0 + wi) : 0xffffffffu;
					}
					out = a.r3_mems + ((uint64_t)(o / kp1) + rd);
					pool12 = a.pool + a.read_pool_off[rd]; n12 = a.read_n_mems[rd];
					nmem = 0; x = 0; ms = me = 0; have = true;
				}
			}
			bool finished = false;
			if (have && !active && !parked) {
				while (x < len && base_at(x) > 3) ++x;
				if (x < len) active = true;
				else { a.r3_n_mems[rd] = nmem; have = false; finished = true; }
			}
			if (!__any_sync(0xffffffffu, finished && !exhausted)) break;
		}
		if (__all_sync(0xffffffffu, !active && !parked)) break;

		// ---- the chain that starts at x ----
		bool walk = false, from_text = false, pre_text = false;     // needs the literal walk / is resolved through the text (pre_text: if its repeat length says so)
		if (!active) ;
		else if (W >= 32 || opt.max_mem_intv < 2) walk = true;      // the shortcuts below assume a 1-row interval ends the chain
		else if (x + W > len || (nmask_window(x) & ((1u << W) - 1u))) {
			// an N or the read end comes before W bases: no seed; the next chain starts after the N (bwt.c:376), or nowhere
			int i = x + 1;
			while (i < len && base_at(i) <= 3) ++i;
			n_ext += (uint32_t)(i - x - 1);
			x = i < len ? i + 1 : len;
		} else if (I.rep && ms <= x && x + W <= me) {
			// W lies inside the one-occurrence SMEM of the previous chain, at text position P: its first K bases occur once iff the
			// longest repeat that starts at P is shorter than K.  One byte at a known place instead of the filter probe -- so the two
			// inverse-SA gathers below do not have to wait for it: all three are issued together (a repeat, rare, throws them away)
			pre_text = true;
		} else {
			const uint32_t c19 = (gather_u32(I.pt + (key_of(x, K) >> 4)) >> (2 * ((uint32_t)key_of(x, K) & 15))) & 3;
			++n_probe;
			if (c19 == 0) { n_ext += (uint32_t)(W - 1); x += W; }   // W does not occur: an x[2] == 0 record, discarded (bwamem.c:260)
			else if (c19 != 1) walk = true;
			else {
				if (!(ms <= x && x + W <= me)) { // find a unique first-pass SMEM around W
					ms = me = 0;
					for (uint32_t m = 0; m < n12; ++m) {
						const uint4 v = reinterpret_cast<const uint4*>(pool12 + m)[1];   // x[2] lo, hi, end, start
						if (v.x == 1 && v.y == 0 && (int)v.w <= x && x + W <= (int)v.z) {
							ms = (int)v.w; me = (int)v.z;
							mtb = gather_u64(I.sa + pool12[m].x[0]); ++n_req;
							break;
						}
					}
				}
				if (me == 0) walk = true;
				else from_text = true;
			}
		}
		// the seed of a chain inside a unique SMEM: rows of W's suffix and of its reverse complement's (all lanes take the
		// LF steps together: the loop is controlled by a vote so that the warp does not drift apart)
		{
			uint64_t r0 = 0, r1 = 0; int w0 = 0, w1 = 0;
			if (from_text || pre_text) {
				const uint64_t P = mtb + (uint64_t)(x - ms);
				uint32_t R = 0;
				if (pre_text) { R = gather_u8(I.rep + P); ++n_req; }
				isa_near(P, r0, w0);
				isa_near(I.seq_len - P - (uint64_t)W, r1, w1);
				if (pre_text) {
					if (R < (uint32_t)K) from_text = true;
					else { walk = true; w0 = w1 = 0; }
				}
			}
			while (__any_sync(0xffffffffu, (w0 | w1) != 0)) {
				if (w0) { r0 = dev_lf(I, r0); --w0; ++n_req; }
				if (w1) { r1 = dev_lf(I, r1); --w1; ++n_req; }
			}
			if (from_text) {
				put(r0, r1, 1, x, x + W);
				n_ext += (uint32_t)(W - 1);
				x += W;
			}
		}
		// ---- literal walk of a chain (bwt.c:366-378), as in k_seed_r3.  It costs ten times a text-assisted chain and few
		//      chains need it, so the lanes that do wait ("parked") until CS_R3_QUORUM of them can walk together, or until
		//      nobody else has anything to do. ----
		if (walk) parked = true;
		{
			const unsigned pm = __ballot_sync(0xffffffffu, parked);
			const unsigned busy = __ballot_sync(0xffffffffu, active && !parked);
			if (!(__popc(pm) >= CS_R3_QUORUM || (pm && !busy))) continue;
		}
		if (parked) {
			parked = false;
			uint64_t c0, c1, c2; int i;
			if (jump >= 2 && x + jump <= len && !(nmask_window(x) & ((1u << jump) - 1u))) {
				kt_lookup(I, (uint32_t)jump, key_of(x, jump), c0, c1, c2);
				i = x + jump; n_ext += (uint32_t)(jump - 1); ++n_req;
			} else {
				const int b = base_at(x);
				c0 = l2_at(I, b) + 1; c1 = l2_at(I, 3 - b) + 1; c2 = l2_at(I, b + 1) - l2_at(I, b);
				i = x + 1;
			}
			for (;;) {
				if (i >= len) { x = len; break; }
				const int b = base_at(i);
				if (b > 3) { x = i + 1; break; }
				++n_ext;
				if (c2 == 0) { // children of an empty interval are empty
					if (i - x >= opt.min_seed_len) { x = i + 1; break; }
					++i; continue;
				}
				uint64_t o0, o1, o2; uint32_t two;
				dev_extend(I, c0, c1, c2, 3 - b, 0, o0, o1, o2, two); ++n_call; n_req += 1u + two;
				if (o2 < (uint64_t)opt.max_mem_intv && i - x >= opt.min_seed_len) {
					if (o2 > 0) put(o0, o1, o2, x, i + 1);
					x = i + 1; break;
				}
				c0 = o0; c1 = o1; c2 = o2; ++i;
			}
		}
	}
	if (n_ext) atomicAdd(a.counters + 0, (unsigned long long)n_ext);
	if (n_call) atomicAdd(a.counters + 1, (unsigned long long)n_call);
	if (n_probe) atomicAdd(a.counters + 3, (unsigned long long)n_probe);
	n_req += n_probe;
	if (n_req) atomicAdd(a.req + 3, (unsigned long long)n_req);
}

// ---------------------------------------------------------------------------------------------
// Collect: put each read's mems in input order, sorted by info (ks_introsort, bwamem.c:271;
// std::sort, comp_seed.cpp:2301), and expand them to SA rows (bwamem.c:386-399).
// One warp per read; rank sort (ties are bit-identical records, so their order is immaterial).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seeds_of(uint64_t x2, int32_t max_occ)
{ // number of iterations of "for (k = count = 0; k < x2 && count < max_occ; k += step, ++count)"
	return x2 < (uint64_t)max_occ ? (uint32_t)x2 : (uint32_t)max_occ;
}

__global__ void k_collect_sort(CollectArgs a)
{ // eight lanes per read (four reads per warp): a read has ~9 mems on ordinary data
	const uint32_t lane = threadIdx.x & 31, sub = lane & 7;
	const unsigned gmask = 0xffu << (lane & 24);
	const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
	const uint32_t kp1 = (uint32_t)a.opt.min_seed_len + 1;
	unsigned long long seeds_total = 0;
	for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < a.n_reads; r += ngroups) {
		const uint32_t n12 = a.read_n_mems[r];                 // passes 1-2 (k_seed_fast, or k_seed alone)
		const uint32_t n3 = a.r3_n_mems ? a.r3_n_mems[r] : 0;  // pass 3 (k_seed_r3)
		const uint32_t n = a.mem_off[r + 1] - a.mem_off[r];    // + the deferred calls of this read (k_seed_walk, k_seed in call mode)
		const cs_mem_t *src12 = a.pool + a.read_pool_off[r];
		const cs_mem_t *src3 = a.r3_mems + ((uint64_t)(a.off[r] / kp1) + r);
		cs_mem_t *dst = a.mems + a.mem_off[r];
		uint32_t n_seeds = 0;
		if ((uint64_t)a.mem_off[r] + n > a.mems_cap) { // result buffer too small: report, never write past it -- but still count the
			auto count = [&](const cs_mem_t *src, uint32_t cnt) { // seeds, so that cs_ctx_need can tell the caller both capacities at once
				for (uint32_t m = sub; m < cnt; m += 8) seeds_total += seeds_of(src[m].x[2], a.opt.max_occ);
			};
			count(src12, n12);
			if (a.read_last_q)
				for (uint32_t q = a.read_last_q[r]; q != 0xffffffffu; q = a.defer_q[q].w) count(a.pool + a.x_off[q], a.x_n[q]);
			count(src3, n3);
			if (sub == 0) { atomicMin(a.error, CS_E_OVERFLOW); a.read_n_seeds[r] = 0; }
			continue;
		}
		const cs_mem_t *all = src12;                           // the read's mems, unsorted, in one place
		if (n != n12) { // several sources: gather them into the staging copy of the output region first
			cs_mem_t *stg = a.stage + a.mem_off[r];
			uint32_t o = 0;
			for (uint32_t m = sub; m < n12; m += 8) { const uint4 *p = reinterpret_cast<const uint4*>(src12 + m); uint4 *d = reinterpret_cast<uint4*>(stg + m); d[0] = p[0]; d[1] = p[1]; }
			o = n12;
			if (a.read_last_q)
				for (uint32_t q = a.read_last_q[r]; q != 0xffffffffu; q = a.defer_q[q].w) {
					const cs_mem_t *sx = a.pool + a.x_off[q];
					const uint32_t nx = a.x_n[q];
					for (uint32_t m = sub; m < nx; m += 8) { const uint4 *p = reinterpret_cast<const uint4*>(sx + m); uint4 *d = reinterpret_cast<uint4*>(stg + o + m); d[0] = p[0]; d[1] = p[1]; }
					o += nx;
				}
			for (uint32_t m = sub; m < n3; m += 8) { const uint4 *p = reinterpret_cast<const uint4*>(src3 + m); uint4 *d = reinterpret_cast<uint4*>(stg + o + m); d[0] = p[0]; d[1] = p[1]; }
			__syncwarp(gmask);
			all = stg;
		}
		for (uint32_t m = sub; m < n; m += 8) {
			const uint4 *p = reinterpret_cast<const uint4*>(all + m);
			uint4 v0 = p[0], v1 = p[1];
			uint64_t info = (uint64_t)v1.z | ((uint64_t)v1.w << 32);
			uint32_t rank = 0;
			for (uint32_t o = 0; o < n; ++o) {
				uint64_t oi = all[o].info;
				rank += (oi < info) || (oi == info && o < m);
			}
			uint4 *d = reinterpret_cast<uint4*>(dst + rank);
			d[0] = v0; d[1] = v1;
			n_seeds += seeds_of((uint64_t)v1.x | ((uint64_t)v1.y << 32), a.opt.max_occ);
		}
		for (int sft = 4; sft > 0; sft >>= 1) n_seeds += __shfl_xor_sync(gmask, n_seeds, sft);
		if (sub == 0) { a.read_n_seeds[r] = n_seeds; seeds_total += n_seeds; }
	}
	for (int sft = 16; sft > 0; sft >>= 1) seeds_total += __shfl_xor_sync(0xffffffffu, seeds_total, sft);
	if (lane == 0 && seeds_total) atomicAdd(a.tot_seeds, seeds_total);
}

// total mems per read (passes 1-2 + pass 3), the input of the offsets scan; the batch totals are also summed in 64 bits
// (tot12: passes 1-2 incl. deferred calls, tot3: pass 3) so that a total past 2^32 -- where the u32 scan would wrap --
// is reported as an overflow instead of passing the capacity checks
__global__ void k_mem_counts(const uint32_t *n12, const uint32_t *n3, const uint32_t *read_last_q, const uint4 *defer_q, const uint32_t *x_n,
                             uint32_t n_reads, uint32_t *out, unsigned long long *tot12, unsigned long long *tot3)
{
	unsigned long long s12 = 0, s3 = 0;
	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
		uint32_t n = n12[r];
		if (read_last_q)
			for (uint32_t q = read_last_q[r]; q != 0xffffffffu; q = defer_q[q].w) n += x_n[q];
		const uint32_t m3 = n3 ? n3[r] : 0;
		s12 += n; s3 += m3;
		out[r] = n + m3;
	}
	for (int sft = 16; sft > 0; sft >>= 1) { s12 += __shfl_xor_sync(0xffffffffu, s12, sft); s3 += __shfl_xor_sync(0xffffffffu, s3, sft); }
	if ((threadIdx.x & 31) == 0) { if (s12) atomicAdd(tot12, s12); if (s3) atomicAdd(tot3, s3); }
}

__global__ void k_collect_rows(CollectArgs a)
{ // eight lanes per read, one mem per lane and trip: a mem has one seed on ordinary data (x[2] == 1), up to max_occ on repeats
	const uint32_t lane = threadIdx.x & 31, sub = lane & 7;
	const unsigned gmask = 0xffu << (lane & 24);
	const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
	const uint64_t max_occ = (uint64_t)a.opt.max_occ;
	for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < a.n_reads; r += ngroups) {
		const uint32_t n = a.mem_off[r + 1] - a.mem_off[r];
		const cs_mem_t *mem = a.mems + a.mem_off[r];
		uint64_t o = a.seed_off[r];
		if ((uint64_t)a.seed_off[r + 1] > a.seed_cap || (uint64_t)a.mem_off[r + 1] > a.mems_cap) {
			if (sub == 0) atomicMin(a.error, CS_E_OVERFLOW);
			continue;
		}
		for (uint32_t m0 = 0; m0 < n; m0 += 8) {
			const uint32_t m = m0 + sub;
			uint64_t x0 = 0, step = 1;
			uint32_t cnt = 0;
			if (m < n) {
				const uint4 v0 = reinterpret_cast<const uint4*>(mem + m)[0], v1 = reinterpret_cast<const uint4*>(mem + m)[1];
				x0 = (uint64_t)v0.x | ((uint64_t)v0.y << 32);
				const uint64_t x2 = (uint64_t)v1.x | ((uint64_t)v1.y << 32);
				cnt = seeds_of(x2, a.opt.max_occ);
				if (x2 > max_occ) step = x2 / max_occ;           // (the division only where the occurrences are sampled, bwamem.c:391)
			}
			uint32_t incl = cnt;                                 // where each lane's seeds start: prefix sum over the eight lanes
#pragma unroll
			for (int d = 1; d < 8; d <<= 1) { const uint32_t v = __shfl_up_sync(gmask, incl, d, 8); if ((int)sub >= d) incl += v; }
			const uint32_t total = __shfl_sync(gmask, incl, 7, 8);
			const uint32_t big = __ballot_sync(gmask, cnt > 4) >> (lane & 24) & 0xffu;
			if (!big) { // the usual case: every lane writes its own few seeds
				for (uint32_t k = 0; k < cnt; ++k) a.seed_rows[o + (incl - cnt) + k] = x0 + (uint64_t)k * step;
			} else { // repeats: the eight lanes share the seeds of one mem after the other
				for (uint32_t j = 0; j < 8 && m0 + j < n; ++j) {
					const uint64_t jx0 = __shfl_sync(gmask, x0, (int)j, 8), jstep = __shfl_sync(gmask, step, (int)j, 8);
					const uint32_t jcnt = __shfl_sync(gmask, cnt, (int)j, 8), jo = __shfl_sync(gmask, incl - cnt, (int)j, 8);
					for (uint32_t k = sub; k < jcnt; k += 8) a.seed_rows[o + jo + k] = jx0 + (uint64_t)k * jstep;
				}
			}
			o += total;
		}
	}
}

// Compact wire format (include/compseed_b200.h, cs_cmem_t): 20 bytes per mem, 5 per seed position.  Runs after
// k_sa_resolve; counts are read on the device.
__global__ void k_compact_results(const uint32_t *n_mems_ptr, uint64_t mems_cap, const cs_mem_t *mems, cs_cmem_t *cmems,
                                  const uint32_t *n_seeds_ptr, uint64_t seeds_cap, const uint64_t *rbeg, uint32_t *lo, uint8_t *hi)
{
	const uint64_t nm = *n_mems_ptr, ns = *n_seeds_ptr;
	const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
	if (nm <= mems_cap)
		for (uint64_t i = tid; i < nm; i += nth) {
			const uint4 *p = reinterpret_cast<const uint4*>(mems + i);
			const uint4 a = p[0], b = p[1];              // x0 lo hi, x1 lo hi | x2 lo hi, end, start
			uint32_t *d = reinterpret_cast<uint32_t*>(cmems + i);
			d[0] = a.x; d[1] = a.z; d[2] = b.x; d[3] = (b.w << 16) | (b.z & 0xffffu);
			d[4] = (a.y & 31u) | ((a.w & 31u) << 5) | ((b.y & 31u) << 10);
		}
	if (ns <= seeds_cap)
		for (uint64_t i = tid; i < ns; i += nth) {
			const uint64_t v = rbeg[i];
			lo[i] = (uint32_t)v; hi[i] = (uint8_t)(v >> 32);
		}
}

// ---------------------------------------------------------------------------------------------
// Measurement helpers
// ---------------------------------------------------------------------------------------------
// Independent uniformly random granule-sized loads over a table: the random-sector roofline.
// `unroll` independent loads are issued back to back per thread before any of them is consumed.
#ifndef CS_EMUL
template <int UNROLL>
__device__ __forceinline__ void gather_body(const uint4 *table, uint64_t n_granules, uint32_t granule16, uint64_t n_loads,
                                            uint64_t seed, unsigned long long *sink)
{
	uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t s = seed ^ (tid * 0x9E3779B97F4A7C15ull);
	uint32_t acc = 0;
	for (uint64_t i = tid; i < n_loads; i += stride * UNROLL) {
		uint64_t a[UNROLL], b[UNROLL], c[UNROLL], d[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			s ^= s << 13; s ^= s >> 7; s ^= s << 17;            // xorshift64
			uint64_t g = (uint64_t)(((unsigned __int128)s * n_granules) >> 64);
			const uint4 *p = table + g * granule16;
			if (granule16 >= 2) {
				asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[u]), "=l"(b[u]), "=l"(c[u]), "=l"(d[u]) : "l"(p));
				if (granule16 >= 4) { // second sector of a 64-byte granule
					uint64_t a2, b2, c2, d2;
					asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a2), "=l"(b2), "=l"(c2), "=l"(d2) : "l"(p + 2));
					a[u] ^= a2; b[u] ^= b2; c[u] ^= c2; d[u] ^= d2;
				}
			} else {
				uint4 v = __ldg(p);
				a[u] = v.x; b[u] = v.y; c[u] = v.z; d[u] = v.w;
			}
		}
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) acc += (uint32_t)(a[u] ^ b[u] ^ c[u] ^ d[u]);
	}
	if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

__global__ void k_gather_probe(const uint4 *table, uint64_t n_granules, uint32_t granule16, uint64_t n_loads, uint64_t seed,
                               unsigned long long *sink)
{ gather_body<1>(table, n_granules, granule16, n_loads, seed, sink); }

__global__ void k_gather_probe4(const uint4 *table, uint64_t n_granules, uint32_t granule16, uint64_t n_loads, uint64_t seed,
                                unsigned long long *sink)
{ gather_body<4>(table, n_granules, granule16, n_loads, seed, sink); }
#endif

__global__ void k_fill(uint4 *p, uint64_t n, uint32_t v)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		p[i] = make_uint4(v, v + 1, v + 2, (uint32_t)i);
}
