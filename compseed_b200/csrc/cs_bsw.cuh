// Device side of the banded Smith-Waterman extension (see cs_bsw.cu for what it reproduces and why it is laid out this way).
// Kept in a header of its own so that tests/emul/bsw_emul.cpp can compile bsw_one_pair as plain C++ and run it on the CPU
// against the reference (test infrastructure; the library has no CPU path).
#pragma once
#include <cstdint>
#include <cstddef>

struct PairIn { int32_t idr, idq, len1, len2, h0; };

struct BswArgs {
	const PairIn *in;
	const uint32_t *order;      // pair indices sorted by query length
	uint32_t n;
	const uint8_t *ref, *qer;
	uint32_t smem_cols;         // k_bsw_extend_smem: columns of a DP row in shared memory (longest query + 2); the packed queries follow the rows
	void *eh; size_t eh_stride; // DP rows, column j of a thread's row at [j * eh_stride + thread] (EhCell: 8 or 4 bytes per cell)
	int32_t w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, max_mat;
	int8_t mat[25];
	int32_t *out;               // 6 per pair: score, tle, gtle, qle, gscore, max_off (the order of SeqPair)
	unsigned int *work;
	unsigned long long *cells;
};

// The DP row of a thread: cell j = {H(i-1, j-1), E(i, j)}.  Both are >= 0 and at most h0 + qlen * max(mat), so when that fits 16 bits
// for every pair of a batch (always, for reads: h0 <= 65535 would need a 64 kbp seed) a cell is ONE 32-bit word, H in the low half.
template <bool WIDE> struct EhCell;
template <> struct EhCell<true>  { typedef int2 T;     static __device__ __forceinline__ T pack(int h, int e) { return make_int2(h, e); }
                                   static __device__ __forceinline__ int h(T v) { return v.x; } static __device__ __forceinline__ int e(T v) { return v.y; }
                                   static __device__ __forceinline__ bool zero(T v) { return (v.x | v.y) == 0; } };
template <> struct EhCell<false> { typedef uint32_t T; static __device__ __forceinline__ T pack(int h, int e) { return (uint32_t)h | ((uint32_t)e << 16); }
                                   static __device__ __forceinline__ int h(T v) { return (int)(v & 0xffffu); } static __device__ __forceinline__ int e(T v) { return (int)(v >> 16); }
                                   static __device__ __forceinline__ bool zero(T v) { return v == 0; } };

// ksw_extend2 (bwalib/ksw.c:380-479) for pair pid; eh: this thread's DP row, column j at eh[j * st]; s_mat: the 5 x 5 scores.
// QPK: the query is first copied, 8 bases per word (4 bits each), to qpk (word k at qpk[k * st]) and read from there: every row of the
// DP re-reads the whole query, and 32 lanes reading one byte each of 32 different queries cost the L1 32 sector look-ups per
// instruction -- more than everything else a cell needs.  Returns the number of cells computed.
template <bool WIDE, bool QPK>
__device__ __forceinline__ unsigned long long bsw_one_pair(const BswArgs &a, uint32_t pid, typename EhCell<WIDE>::T *eh, size_t st, const int *s_mat, uint32_t *qpk = nullptr)
{
	typedef EhCell<WIDE> Cell;
	typedef typename Cell::T cell_t;
	const int o_del = a.o_del, e_del = a.e_del, o_ins = a.o_ins, e_ins = a.e_ins, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, zdrop = a.zdrop;
	unsigned long long cells = 0;
	const PairIn p = a.in[pid];
	const int qlen = p.len2, tlen = p.len1, h0 = p.h0;
	const uint8_t *query = a.qer + p.idq, *target = a.ref + p.idr;
	int i, j, beg, end, max, max_i, max_j, max_ie, gscore, max_off, w = a.w;
	if (QPK)
		for (j = 0; j < qlen; j += 8) {
			uint32_t v = 0;
			for (int k = 0; k < 8 && j + k < qlen; ++k) v |= (uint32_t)(query[j + k] & 15) << (4 * k);
			qpk[(size_t)(j >> 3) * st] = v;
		}
	// first row (ksw.c:395-398): H(-1, j); everything else of the row array is zero (calloc)
	{
		int h = h0 > oe_ins ? h0 - oe_ins : 0;
		eh[0] = Cell::pack(h0, 0);
		eh[st] = Cell::pack(h, 0);
		for (j = 2; j <= qlen; ++j) {
			h = h > e_ins ? h - e_ins : 0;      // (once 0 <= e_ins the reference's loop stops and the rest stays 0)
			eh[(size_t)j * st] = Cell::pack(h, 0);
		}
	}
	// w no larger than what the scores allow (ksw.c:401-408)
	{
		int max_ins = (int)((double)(qlen * a.max_mat + a.end_bonus - o_ins) / e_ins + 1.);
		max_ins = max_ins > 1 ? max_ins : 1;
		w = w < max_ins ? w : max_ins;
		int max_del = (int)((double)(qlen * a.max_mat + a.end_bonus - o_del) / e_del + 1.);
		max_del = max_del > 1 ? max_del : 1;
		w = w < max_del ? w : max_del;
	}
	max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
	beg = 0; end = qlen;
	for (i = 0; i < tlen; ++i) {
		int f = 0, h1, m = 0, mj = -1;
		const int *q = s_mat + 5 * (int)target[i];
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
		else h1 = 0;
		cell_t *pe = eh + (size_t)beg * st;
		// (fetching the next column's cell ahead of this column's store was tried: 4 x slower, profiles/r02_bsw_variants.txt)
		for (j = beg; j < end; ++j, pe += st) { // ksw.c:421-447
			const cell_t v = *pe;
			int M = Cell::h(v), e = Cell::e(v), h, t;
			M = M ? M + q[QPK ? (int)((qpk[(size_t)(j >> 3) * st] >> (4 * (j & 7))) & 15u) : (int)query[j]] : 0;
			h = M > e ? M : e;
			h = h > f ? h : f;
			mj = m > h ? mj : j;
			m = m > h ? m : h;
			t = M - oe_del; t = t > 0 ? t : 0;
			e -= e_del; e = e > t ? e : t;
			*pe = Cell::pack(h1, e);
			h1 = h;
			t = M - oe_ins; t = t > 0 ? t : 0;
			f -= e_ins; f = f > t ? f : t;
		}
		cells += (unsigned)(end > beg ? end - beg : 0);
		eh[(size_t)end * st] = Cell::pack(h1, 0);   // (end, not where the loop stopped: the band can be empty, beg > end, once i - w passes the query)
		if (j == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
		if (m == 0) break;
		if (m > max) {
			max = m; max_i = i; max_j = mj;
			const int off = mj > i ? mj - i : i - mj;
			max_off = max_off > off ? max_off : off;
		} else if (zdrop > 0) {
			if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
			else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
		}
		// the band of the next row: the non-zero cells of this one (ksw.c:463-468)
		for (j = beg; j < end; ++j) if (!Cell::zero(eh[(size_t)j * st])) break;
		beg = j;
		for (j = end; j >= beg; --j) if (!Cell::zero(eh[(size_t)j * st])) break;
		end = j + 2 < qlen ? j + 2 : qlen;
	}
	int32_t *o = a.out + 6 * (size_t)pid;
	o[0] = max; o[1] = max_i + 1; o[2] = max_ie + 1; o[3] = max_j + 1; o[4] = gscore; o[5] = max_off;
	return cells;
}

// The same with the DP rows in SHARED memory ([column][thread of the CTA], bank-conflict free): the row loads of k_bsw_extend below go
// to an interleaved HBM scratch that only partly stays in L2 and leave the warps waiting on the long scoreboard (ncu: 18 % issue
// active, 34 stalled warps per issue, profiles/r02_ncu_k_bsw_extend_2Mpairs.txt); a row of <= a few hundred cells per thread fits
// shared memory for about ten warps per SM, and at shared-memory latency that many are enough.  Used when the longest query of
// the batch allows at least two CTAs per SM; k_bsw_extend is the fallback for longer queries.
#ifndef CS_BSW_SMEM_BLOCK
#define CS_BSW_SMEM_BLOCK 32   // one warp per CTA: the finest granularity for shared memory (2 M pairs: 32: 412 GCUPS, 64: 399, 128: 334; profiles/r02_bsw_variants.txt)
#endif
#ifndef CS_BSW_EMUL   // (tests/emul/bsw_emul.cpp runs bsw_one_pair alone)
template <bool WIDE>
__global__ void __launch_bounds__(CS_BSW_SMEM_BLOCK) k_bsw_extend_smem(BswArgs a)
{
	extern __shared__ uint4 s_rows_raw[];
	__shared__ int s_mat[25];
	if (threadIdx.x < 25) s_mat[threadIdx.x] = a.mat[threadIdx.x];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	typename EhCell<WIDE>::T *const eh = reinterpret_cast<typename EhCell<WIDE>::T*>(s_rows_raw) + threadIdx.x;
	uint32_t *const qpk = reinterpret_cast<uint32_t*>(reinterpret_cast<typename EhCell<WIDE>::T*>(s_rows_raw) + (size_t)CS_BSW_SMEM_BLOCK * a.smem_cols) + threadIdx.x;
	unsigned long long cells = 0;
	for (;;) { // 32 pairs of similar shape per warp and trip
		unsigned int base = 0;
		if (lane == 0) base = atomicAdd(a.work, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= a.n) break;
		if (base + lane < a.n) cells += bsw_one_pair<WIDE, true>(a, a.order[a.n - 1 - (base + lane)], eh, CS_BSW_SMEM_BLOCK, s_mat, qpk);   // largest shapes first
	}
	if (cells) atomicAdd(a.cells, cells);
}
#endif

template <bool WIDE>
__global__ void __launch_bounds__(128) k_bsw_extend(BswArgs a)
{
	__shared__ int s_mat[25];
	if (threadIdx.x < 25) s_mat[threadIdx.x] = a.mat[threadIdx.x];
	__syncthreads();
	const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	typename EhCell<WIDE>::T *const eh = reinterpret_cast<typename EhCell<WIDE>::T*>(a.eh) + gtid;
	unsigned long long cells = 0;
	for (;;) { // 32 pairs of similar shape per warp and trip
		unsigned int base = 0;
		if (lane == 0) base = atomicAdd(a.work, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= a.n) break;
		if (base + lane < a.n) cells += bsw_one_pair<WIDE, false>(a, a.order[a.n - 1 - (base + lane)], eh, a.eh_stride, s_mat);   // largest shapes first
	}
	if (cells) atomicAdd(a.cells, cells);
}
