// TEST INFRASTRUCTURE ONLY.  Compiles the device function of compseed_b200/csrc/cs_bsw.cuh (bsw_one_pair: one ksw_extend2 per
// sequence pair) as plain C++, the CUDA qualifiers defined away, and runs it serially on the CPU, so that the extension kernel's
// source can be checked against the reference's BandedPairWiseSW / ksw_extend2 without a GPU (tests/test_bsw_emul.py).
// The shipped library never contains or runs this build: there is no CPU path in the product.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <vector>

#define CS_BSW_EMUL 1
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
struct emul_dim3 { unsigned x, y, z; };
static emul_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline void __syncthreads() {}
static inline unsigned __shfl_sync(unsigned, unsigned v, int) { return v; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }

#include "../../compseed_b200/csrc/cs_bsw.cuh"

extern "C" {

// pairs: 14 int32 each (SeqPair); eh_stride > 1 exercises the interleaved row layout of the kernel.  Returns the cells computed.
unsigned long long bsw_emul(int32_t *pairs, const uint8_t *ref, const uint8_t *qer, uint32_t n, int w, int o_del, int e_del, int o_ins, int e_ins,
                            int zdrop, int end_bonus, const int8_t *mat, uint32_t eh_stride, int wide, int qpack)
{
	std::vector<PairIn> in(n);
	std::vector<int32_t> out((size_t)n * 6);
	int max_q = 1;
	for (uint32_t i = 0; i < n; ++i) {
		const int32_t *p = pairs + 14 * (size_t)i;
		in[i].idr = p[0]; in[i].idq = p[1]; in[i].len1 = p[3]; in[i].len2 = p[4]; in[i].h0 = p[5];
		if (p[4] > max_q) max_q = p[4];
	}
	std::vector<int2> eh((size_t)(max_q + 1) * eh_stride);
	for (auto &v : eh) v = make_int2(0x5a5a5a5a, 0x5a5a5a5a);   // stale rows of earlier pairs must not matter
	BswArgs a;
	memset(&a, 0, sizeof a);
	a.in = in.data(); a.n = n; a.ref = ref; a.qer = qer; a.w = w; a.o_del = o_del; a.e_del = e_del; a.o_ins = o_ins; a.e_ins = e_ins;
	a.zdrop = zdrop; a.end_bonus = end_bonus; a.out = out.data();
	int s_mat[25], mx = 0;
	for (int k = 0; k < 25; ++k) { a.mat[k] = mat[k]; s_mat[k] = mat[k]; mx = mx > mat[k] ? mx : mat[k]; }
	a.max_mat = mx;
	unsigned long long cells = 0;
	std::vector<uint32_t> qpk((size_t)(max_q / 8 + 2) * eh_stride, 0xdeadbeefu);
	if (qpack) { // the variant the shared-memory kernel runs: query read from its 4-bit packed copy
		for (uint32_t i = 0; i < n; ++i)
			cells += wide ? bsw_one_pair<true, true>(a, i, eh.data() + (i % eh_stride), eh_stride, s_mat, qpk.data() + (i % eh_stride))
			              : bsw_one_pair<false, true>(a, i, reinterpret_cast<uint32_t*>(eh.data()) + (i % eh_stride), eh_stride, s_mat, qpk.data() + (i % eh_stride));
	} else
	for (uint32_t i = 0; i < n; ++i)
		cells += wide ? bsw_one_pair<true, false>(a, i, eh.data() + (i % eh_stride), eh_stride, s_mat)
		              : bsw_one_pair<false, false>(a, i, reinterpret_cast<uint32_t*>(eh.data()) + (i % eh_stride), eh_stride, s_mat);
	for (uint32_t i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) pairs[14 * (size_t)i + 8 + k] = out[6 * (size_t)i + k];
	return cells;
}

}
