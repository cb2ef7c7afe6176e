// Banded Smith-Waterman extension on the device: one ksw_extend2 (bwalib/ksw.c:380-479 ==
// BandedPairWiseSW::scalarBandedSWA, mapping/bandedSWA.cpp:118-237) per sequence pair, behind the batch interface of the
// reference's extension stage (scalarBandedSWAWrapper / getScores8 / getScores16, called from mem_chain2aln_across_reads_V2,
// mapping/comp_seed.cpp:1722-2074).
//
// What has to be reproduced bit for bit, besides the recurrences: (1) the band is not the fixed |i - j| <= w diagonal band -- after
// every row it is cut back to the cells that are non-zero (H or E) plus one column to the right (ksw.c:463-468), so an insertion
// run that would cross the cut is lost, and rows end early; (2) the row maximum remembers the LAST column that reaches it
// (ksw.c:438), the global maximum the FIRST row (:452); (3) the z-drop test compares the diagonal offsets of the two maxima with
// separate insertion / deletion penalties (:456-462); (4) w is capped by what the scores allow, computed in double (:401-408);
// (5) F (a gap in the target) is opened from M, the diagonal move, not from H (:429,445) -- which makes F of a row a plain
// prefix maximum over that row's M values, and E depend on the row above only.
//
// Mapping.  One pair per THREAD: the work of a pair is a chain of dependent rows whose width changes from row to row, typically a
// few tens of cells; a warp per pair would leave most lanes idle.  The DP row (H(i-1,j-1), E(i,j)) of a thread lives in a scratch
// array interleaved over the threads ([column][thread]), so that the lanes of a warp, which sweep their rows in lock step, touch
// consecutive words.  Pairs are handed out 32 at a time from a list sorted by query length (the reference sorts its pairs by
// length too, sortPairsLenExt, for the same reason: lanes of one vector should finish together).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>
#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include "cs_internal.h"

#include "cs_bsw.cuh"

namespace {

// sort key: query length (the width of a row), then target length (the number of rows) in steps of 8 -- the 32 pairs of a warp
// should have the same shape
#define BSW_KEY_BITS 26
__global__ void k_bsw_keys(const PairIn *in, uint32_t n, uint32_t *keys, uint32_t *idx)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t t = (uint32_t)in[i].len1 >> 3;
#ifdef BSW_KEY_QLEN_ONLY
		keys[i] = (uint32_t)in[i].len2 << 9; idx[i] = i; (void)t;
#else
		keys[i] = ((uint32_t)in[i].len2 << 9) | (t > 511u ? 511u : t); idx[i] = i;
#endif
	}
}

} // namespace

// Host loops over the pairs of a batch (conversion in, scores out) and the copies of the two sequence buffers into page-locked
// memory, split over the host's threads: at 50 M pairs/s on the device one thread doing them would be the bottleneck.
template <class F> static void host_par_for(uint64_t n, uint64_t grain, F fn)
{
	unsigned nt = std::thread::hardware_concurrency();
	if (nt > 16) nt = 16;
	if (nt < 2 || n < 2 * grain) { fn((uint64_t)0, n); return; }
	if ((uint64_t)nt > n / grain) nt = (unsigned)(n / grain);
	std::vector<std::thread> th;
	const uint64_t per = (n + nt - 1) / nt;
	for (unsigned k = 0; k < nt; ++k) {
		const uint64_t lo = (uint64_t)k * per, hi = std::min<uint64_t>(n, lo + per);
		if (lo < hi) th.emplace_back([=]() { fn(lo, hi); });
	}
	for (auto &t : th) t.join();
}

struct cs_bsw {
	int device, n_sm;
	cudaStream_t stream;
	cudaEvent_t ev0, ev1;
	uint32_t cap_pairs, n_pairs, max_qlen_staged;
	int64_t max_h0_staged;   // of the staged batch: decides whether a DP cell fits 32 bits (EhCell)
	bool neg_h0_staged;
	int ctas_per_sm;         // resident 128-thread CTAs of the extension kernel per SM (cs_bsw_set_ctas_per_sm)
	int use_smem;            // -1 / 1: DP rows in shared memory when the longest query allows (k_bsw_extend_smem); 0: always the HBM scratch
	uint64_t cap_ref, cap_qer;
	PairIn *h_in, *d_in;
	int32_t *h_out, *d_out;
	uint8_t *h_ref, *h_qer, *d_ref, *d_qer;
	uint32_t *d_keys, *d_keys2, *d_idx, *d_order;
	void *d_sort_tmp; size_t sort_tmp_bytes;
	int2 *d_eh; size_t eh_threads; uint32_t eh_qlen;   // scratch: eh_threads x (eh_qlen + 1)
	unsigned int *d_work; unsigned long long *d_cells;
	unsigned long long h_cells;
	uint64_t n_launch;
	bool staged, ran;
};

static void bsw_free_pairs(cs_bsw *b)
{
	cudaFreeHost(b->h_in); cudaFreeHost(b->h_out); cudaFree(b->d_in); cudaFree(b->d_out);
	cudaFree(b->d_keys); cudaFree(b->d_keys2); cudaFree(b->d_idx); cudaFree(b->d_order); cudaFree(b->d_sort_tmp);
	b->h_in = nullptr; b->h_out = nullptr; b->d_in = nullptr; b->d_out = nullptr; b->d_keys = b->d_keys2 = b->d_idx = b->d_order = nullptr; b->d_sort_tmp = nullptr;
}

static int bsw_alloc_pairs(cs_bsw *b, uint32_t cap)
{
	bsw_free_pairs(b);
	b->cap_pairs = 0;
	CK(cudaMallocHost(&b->h_in, (size_t)cap * sizeof(PairIn))); CK(cudaMallocHost(&b->h_out, (size_t)cap * 24));
	CK(cudaMalloc(&b->d_in, (size_t)cap * sizeof(PairIn))); CK(cudaMalloc(&b->d_out, (size_t)cap * 24));
	CK(cudaMalloc(&b->d_keys, (size_t)cap * 4)); CK(cudaMalloc(&b->d_keys2, (size_t)cap * 4)); CK(cudaMalloc(&b->d_idx, (size_t)cap * 4)); CK(cudaMalloc(&b->d_order, (size_t)cap * 4));
	b->sort_tmp_bytes = 0;
	CK(cub::DeviceRadixSort::SortPairs(nullptr, b->sort_tmp_bytes, b->d_keys, b->d_keys2, b->d_idx, b->d_order, (int)cap, 0, BSW_KEY_BITS, b->stream));
	CK(cudaMalloc(&b->d_sort_tmp, b->sort_tmp_bytes + 256));
	b->cap_pairs = cap;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

static int bsw_alloc_seq(uint8_t **h, uint8_t **d, uint64_t *cap, uint64_t need)
{
	cudaFreeHost(*h); cudaFree(*d); *h = nullptr; *d = nullptr; *cap = 0;
	CK(cudaMallocHost(h, need + 16)); CK(cudaMalloc(d, need + 16));
	*cap = need;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

extern "C" void cs_bsw_free(cs_bsw_t *b)
{
	if (!b) return;
	cudaSetDevice(b->device);
	bsw_free_pairs(b);
	cudaFreeHost(b->h_ref); cudaFreeHost(b->h_qer); cudaFree(b->d_ref); cudaFree(b->d_qer); cudaFree(b->d_eh); cudaFree(b->d_work); cudaFree(b->d_cells);
	if (b->ev0) cudaEventDestroy(b->ev0);
	if (b->ev1) cudaEventDestroy(b->ev1);
	if (b->stream) cudaStreamDestroy(b->stream);
	free(b);
}

extern "C" cs_bsw_t *cs_bsw_create(int device, uint32_t max_pairs, uint64_t max_ref_bytes, uint64_t max_qer_bytes, uint32_t max_qlen)
{
	if (max_pairs == 0 || max_pairs >= (1u << 31)) { cs_set_err(CS_E_ARG, "max_pairs %u outside [1, 2^31)", max_pairs); return nullptr; }
	if (cs_use_device(device) != CS_OK) return nullptr;
	cs_bsw *b = (cs_bsw*)calloc(1, sizeof(cs_bsw));
	if (!b) { cs_set_err(CS_E_ARG, "out of host memory"); return nullptr; }
	b->device = device;
	{
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		b->n_sm = prop.multiProcessorCount;
	}
	CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	CK(cudaEventCreate(&b->ev0)); CK(cudaEventCreate(&b->ev1));
	CK(cudaMalloc(&b->d_work, 4)); CK(cudaMalloc(&b->d_cells, 8));
	// the DP rows live in L1: all of it (the kernel uses 100 bytes of shared memory)
	CK(cudaFuncSetAttribute(k_bsw_extend<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
	CK(cudaFuncSetAttribute(k_bsw_extend<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
	if (bsw_alloc_pairs(b, max_pairs) != CS_OK) goto fail;
	if (bsw_alloc_seq(&b->h_ref, &b->d_ref, &b->cap_ref, std::max<uint64_t>(max_ref_bytes, 4096)) != CS_OK) goto fail;
	if (bsw_alloc_seq(&b->h_qer, &b->d_qer, &b->cap_qer, std::max<uint64_t>(max_qer_bytes, 4096)) != CS_OK) goto fail;
	b->eh_qlen = 0; b->eh_threads = 0; b->max_qlen_staged = max_qlen; b->ctas_per_sm = 8; b->use_smem = -1;
	return b;
fail:
	cs_bsw_free(b);
	return nullptr;
}

extern "C" uint64_t cs_bsw_launches(const cs_bsw_t *b) { return b ? b->n_launch : 0; }

extern "C" int cs_bsw_set_rows_in_smem(cs_bsw_t *b, int on)
{
	if (!b) return cs_set_err(CS_E_ARG, "null argument");
	b->use_smem = on ? 1 : 0;
	return CS_OK;
}

extern "C" int cs_bsw_set_ctas_per_sm(cs_bsw_t *b, int ctas_per_sm)
{
	if (!b || ctas_per_sm < 1 || ctas_per_sm > 16) return cs_set_err(CS_E_ARG, "ctas_per_sm outside [1, 16]");
	b->ctas_per_sm = ctas_per_sm;
	return CS_OK;
}

extern "C" int cs_bsw_stage(cs_bsw_t *b, const cs_seqpair_t *pairs, const uint8_t *seq_buf_ref, uint64_t ref_bytes, const uint8_t *seq_buf_qer, uint64_t qer_bytes, uint32_t n_pairs)
{
	if (!b || !pairs || !seq_buf_ref || !seq_buf_qer) return cs_set_err(CS_E_ARG, "null argument");
	if (n_pairs == 0 || n_pairs >= (1u << 31)) return cs_set_err(CS_E_ARG, "n_pairs %u outside [1, 2^31)", n_pairs);
	{ const int rc = cs_use_device(b->device); if (rc != CS_OK) return rc; }
	b->staged = b->ran = false;
	if (n_pairs > b->cap_pairs && bsw_alloc_pairs(b, n_pairs + n_pairs / 4) != CS_OK) return CS_E_CUDA;
	if (ref_bytes > b->cap_ref && bsw_alloc_seq(&b->h_ref, &b->d_ref, &b->cap_ref, ref_bytes + ref_bytes / 4) != CS_OK) return CS_E_CUDA;
	if (qer_bytes > b->cap_qer && bsw_alloc_seq(&b->h_qer, &b->d_qer, &b->cap_qer, qer_bytes + qer_bytes / 4) != CS_OK) return CS_E_CUDA;
	uint32_t mq = 1; int64_t mh = 0; bool neg = false;
	{
		std::atomic<uint32_t> a_mq(1), a_bad(0xffffffffu);
		std::atomic<long long> a_mh(0);
		std::atomic<bool> a_neg(false);
		PairIn *h_in = b->h_in;
		host_par_for(n_pairs, 1u << 16, [&, h_in](uint64_t lo, uint64_t hi) {
			uint32_t lmq = 1, lbad = 0xffffffffu; long long lmh = 0; bool lneg = false;
			for (uint64_t i = lo; i < hi; ++i) {
				const cs_seqpair_t &p = pairs[i];
				if (p.len2 < 1 || p.len1 < 0 || p.idr < 0 || p.idq < 0 || (uint64_t)p.idr + (uint64_t)p.len1 > ref_bytes || (uint64_t)p.idq + (uint64_t)p.len2 > qer_bytes) {
					if ((uint32_t)i < lbad) lbad = (uint32_t)i;
					continue;
				}
				h_in[i].idr = p.idr; h_in[i].idq = p.idq; h_in[i].len1 = p.len1; h_in[i].len2 = p.len2; h_in[i].h0 = p.h0;
				if ((uint32_t)p.len2 > lmq) lmq = (uint32_t)p.len2;
				if (p.h0 > lmh) lmh = p.h0;
				if (p.h0 < 0) lneg = true;
			}
			uint32_t cur = a_mq.load(); while (lmq > cur && !a_mq.compare_exchange_weak(cur, lmq)) {}
			long long ch = a_mh.load(); while (lmh > ch && !a_mh.compare_exchange_weak(ch, lmh)) {}
			uint32_t cb = a_bad.load(); while (lbad < cb && !a_bad.compare_exchange_weak(cb, lbad)) {}
			if (lneg) a_neg.store(true);
		});
		if (a_bad.load() != 0xffffffffu) {
			const uint32_t i = a_bad.load();
			const cs_seqpair_t &p = pairs[i];
			return cs_set_err(CS_E_ARG, "pair %u: idr %d len1 %d / idq %d len2 %d do not fit the buffers (%llu, %llu bytes; a query has at least one base)",
			                  i, p.idr, p.len1, p.idq, p.len2, (unsigned long long)ref_bytes, (unsigned long long)qer_bytes);
		}
		mq = a_mq.load(); mh = a_mh.load(); neg = a_neg.load();
	}
	if (mq >= (1u << 17)) return cs_set_err(CS_E_ARG, "query of %u bases: longer than a read can be (65535, comp_seed.h:39)", mq);
	b->max_qlen_staged = mq; b->max_h0_staged = mh; b->neg_h0_staged = neg;
	{
		auto pinned = [](const void *p) { cudaPointerAttributes pa; const bool ok = cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeHost; cudaGetLastError(); return ok; };
		const uint8_t *sr = seq_buf_ref, *sq = seq_buf_qer;
		if (!pinned(sr)) { uint8_t *dst = b->h_ref; host_par_for(ref_bytes, 4u << 20, [=](uint64_t lo, uint64_t hi) { memcpy(dst + lo, sr + lo, hi - lo); }); sr = b->h_ref; }
		if (!pinned(sq)) { uint8_t *dst = b->h_qer; host_par_for(qer_bytes, 4u << 20, [=](uint64_t lo, uint64_t hi) { memcpy(dst + lo, sq + lo, hi - lo); }); sq = b->h_qer; }
		CK(cudaMemcpyAsync(b->d_in, b->h_in, (size_t)n_pairs * sizeof(PairIn), cudaMemcpyHostToDevice, b->stream));
		CK(cudaMemcpyAsync(b->d_ref, sr, ref_bytes, cudaMemcpyHostToDevice, b->stream));
		CK(cudaMemcpyAsync(b->d_qer, sq, qer_bytes, cudaMemcpyHostToDevice, b->stream));
		CK(cudaStreamSynchronize(b->stream));    // (the caller's buffers are free again)
	}
	b->n_pairs = n_pairs; b->staged = true;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

extern "C" int cs_bsw_run_staged(cs_bsw_t *b, int32_t w, const cs_bsw_opt_t *opt, float *kernel_ms, uint64_t *cells)
{
	if (!b || !opt) return cs_set_err(CS_E_ARG, "null argument");
	if (!b->staged) return cs_set_err(CS_E_STATE, "no staged pairs (cs_bsw_stage)");
	if (w < 0 || opt->e_del < 1 || opt->e_ins < 1 || opt->o_del < 0 || opt->o_ins < 0)
		return cs_set_err(CS_E_ARG, "bad extension options (w %d, o_del %d, e_del %d, o_ins %d, e_ins %d)", w, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins);
	{ const int rc = cs_use_device(b->device); if (rc != CS_OK) return rc; }
	const uint32_t n = b->n_pairs;
	// DP rows: one per resident thread, max_qlen + 1 columns; fewer threads for very long queries (2 GiB of scratch at most)
	int grid = b->n_sm * b->ctas_per_sm;
	{
		const uint64_t cols = (uint64_t)b->max_qlen_staged + 1;
		while (grid > b->n_sm && (uint64_t)grid * 128 * cols * 8 > (2ull << 30)) grid -= b->n_sm;
		while (grid > 1 && (uint64_t)grid * 128 * cols * 8 > (2ull << 30)) grid /= 2;
		grid = std::min<int>(grid, (int)((n + 127) / 128));
		if (grid < 1) grid = 1;
		const size_t threads = (size_t)grid * 128;
		if (threads > b->eh_threads || b->max_qlen_staged > b->eh_qlen) {
			const size_t nt = std::max(threads, b->eh_threads); const uint32_t nq = std::max(b->max_qlen_staged, b->eh_qlen);
			cudaFree(b->d_eh); b->d_eh = nullptr; b->eh_threads = 0; b->eh_qlen = 0;
			CK(cudaMalloc(&b->d_eh, nt * ((size_t)nq + 1) * sizeof(int2)));
			b->eh_threads = nt; b->eh_qlen = nq;
		}
	}
	{
		BswArgs a;
		a.in = b->d_in; a.order = b->d_order; a.n = n; a.ref = b->d_ref; a.qer = b->d_qer;
		a.eh = b->d_eh; a.eh_stride = b->eh_threads;
		a.w = w; a.o_del = opt->o_del; a.e_del = opt->e_del; a.o_ins = opt->o_ins; a.e_ins = opt->e_ins; a.zdrop = opt->zdrop; a.end_bonus = opt->end_bonus;
		int mx = 0;
		for (int k = 0; k < 25; ++k) { a.mat[k] = opt->mat[k]; mx = mx > opt->mat[k] ? mx : opt->mat[k]; }
		a.max_mat = mx;
		a.out = b->d_out; a.work = b->d_work; a.cells = b->d_cells;
		CK(cudaMemsetAsync(b->d_work, 0, 4, b->stream)); CK(cudaMemsetAsync(b->d_cells, 0, 8, b->stream));
		k_bsw_keys<<<std::min<int>(b->n_sm * 8, (int)((n + 255) / 256)), 256, 0, b->stream>>>(b->d_in, n, b->d_keys, b->d_idx);
		CK(cudaGetLastError()); ++b->n_launch;
		CK(cub::DeviceRadixSort::SortPairs(b->d_sort_tmp, b->sort_tmp_bytes, b->d_keys, b->d_keys2, b->d_idx, b->d_order, (int)n, 0, BSW_KEY_BITS, b->stream));
		b->n_launch += 5;   // cub: histogram + onesweep passes of a 26-bit key
		CK(cudaEventRecord(b->ev0, b->stream));
		// a cell holds two values in [0, h0 + qlen * max(mat)]: 16 bits each when that allows (negative h0 never does)
		{
#ifdef BSW_FORCE_WIDE
			const bool wide = true;
#else
			const bool wide = b->neg_h0_staged || b->max_h0_staged + (int64_t)b->max_qlen_staged * (mx > 0 ? mx : 0) >= 65536;
#endif
			// rows in shared memory when two or more CTAs fit an SM with the longest query of the batch (b->use_smem: -1 auto)
			const size_t smem = (size_t)CS_BSW_SMEM_BLOCK * (((size_t)b->max_qlen_staged + 2) * (wide ? 8 : 4) + ((size_t)b->max_qlen_staged / 8 + 2) * 4);
			a.smem_cols = b->max_qlen_staged + 2;
			int per_sm = 0;
			if (b->use_smem != 0 && smem <= 100 * 1024) {
				if (wide) { CK(cudaFuncSetAttribute(k_bsw_extend_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bsw_extend_smem<true>, CS_BSW_SMEM_BLOCK, smem)); }
				else { CK(cudaFuncSetAttribute(k_bsw_extend_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				       CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bsw_extend_smem<false>, CS_BSW_SMEM_BLOCK, smem)); }
			}
			if (per_sm >= 2) {
				const int g = std::max(1, std::min<int>(b->n_sm * per_sm, (int)((n + CS_BSW_SMEM_BLOCK - 1) / CS_BSW_SMEM_BLOCK)));
				if (wide) k_bsw_extend_smem<true><<<g, CS_BSW_SMEM_BLOCK, smem, b->stream>>>(a);
				else k_bsw_extend_smem<false><<<g, CS_BSW_SMEM_BLOCK, smem, b->stream>>>(a);
			} else if (wide) k_bsw_extend<true><<<grid, 128, 0, b->stream>>>(a);
			else k_bsw_extend<false><<<grid, 128, 0, b->stream>>>(a);
		}
		CK(cudaGetLastError()); ++b->n_launch;
		CK(cudaEventRecord(b->ev1, b->stream));
		CK(cudaMemcpyAsync(&b->h_cells, b->d_cells, 8, cudaMemcpyDeviceToHost, b->stream));
		CK(cudaStreamSynchronize(b->stream));
	}
	if (kernel_ms) cudaEventElapsedTime(kernel_ms, b->ev0, b->ev1);
	if (cells) *cells = b->h_cells;
	b->ran = true;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

extern "C" int cs_bsw_fetch(cs_bsw_t *b, cs_seqpair_t *pairs)
{
	if (!b || !pairs) return cs_set_err(CS_E_ARG, "null argument");
	if (!b->ran) return cs_set_err(CS_E_STATE, "no finished run (cs_bsw_run_staged)");
	{ const int rc = cs_use_device(b->device); if (rc != CS_OK) return rc; }
	CK(cudaMemcpyAsync(b->h_out, b->d_out, (size_t)b->n_pairs * 24, cudaMemcpyDeviceToHost, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	{
		const int32_t *h_out = b->h_out;
		host_par_for(b->n_pairs, 1u << 16, [=](uint64_t lo, uint64_t hi) {
			for (uint64_t i = lo; i < hi; ++i) {
				const int32_t *o = h_out + 6 * (size_t)i;
				pairs[i].score = o[0]; pairs[i].tle = o[1]; pairs[i].gtle = o[2]; pairs[i].qle = o[3]; pairs[i].gscore = o[4]; pairs[i].max_off = o[5];
			}
		});
	}
	return CS_OK;
fail:
	return CS_E_CUDA;
}

extern "C" int cs_bsw_extend(cs_bsw_t *b, cs_seqpair_t *pairs, const uint8_t *seq_buf_ref, uint64_t ref_bytes, const uint8_t *seq_buf_qer, uint64_t qer_bytes,
                             uint32_t n_pairs, int32_t w, const cs_bsw_opt_t *opt)
{
	int rc;
	if (n_pairs == 0) return CS_OK;   // (the reference's wrapper loops over zero pairs)
	if ((rc = cs_bsw_stage(b, pairs, seq_buf_ref, ref_bytes, seq_buf_qer, qer_bytes, n_pairs)) != CS_OK) return rc;
	if ((rc = cs_bsw_run_staged(b, w, opt, nullptr, nullptr)) != CS_OK) return rc;
	return cs_bsw_fetch(b, pairs);
}
