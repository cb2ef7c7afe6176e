// TEST INFRASTRUCTURE ONLY.  Compiles compseed_b200/csrc/cs_chain.cu as plain C++ (the CUDA qualifiers defined away, the
// kernels' thread indices and atomics stubbed) and runs its kernels serially on the CPU, so that the chaining stage can
// be checked against the reference's mem_chain / mem_chain_flt without a GPU (tests/test_chain_emul.py).  The shipped
// library never contains or runs this build: there is no CPU path in the product.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
struct emul_dim3 { unsigned x, y, z; };
static emul_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }

#include "../../compseed_b200/csrc/cs_chain.cu"

extern "C" {

// Same inputs as the device stage gets from a batch; outputs sized by the caller (n_chains <= n_seeds, chain seeds <= n_seeds).
// Returns 0, or the error code the kernels flagged.
int chain_emul(uint32_t n_reads, const uint32_t *read_off, const uint32_t *mem_off, const cs_mem_t *mems, const uint32_t *seed_off, const int64_t *rbeg,
               const cs_seed_opt_t *opt, const cs_chain_opt_t *copt, int64_t l_pac, int32_t n_seqs, const int64_t *c_off, const uint8_t *c_alt,
               uint32_t *chain_off, uint32_t *cseed_off, cs_chain_t *chains, uint32_t *s_lo, uint8_t *s_hi, uint16_t *s_qbeg, uint16_t *s_len)
{
	const uint64_t ns = seed_off[n_reads];
	std::vector<uint32_t> s_next(ns + 1), s_qb_len(ns + 1), order(ns + 1), klist(ns + 1), node_cnt(n_reads + 1), node_off(n_reads + 1);
	std::vector<ChainTmp> tmp(ns + 1);
	std::vector<uint32_t> n_chain(n_reads + 1), n_cseed(n_reads + 1), l_rep(n_reads + 1), n_seeds(n_reads + 1);
	for (uint32_t r = 0; r < n_reads; ++r) n_seeds[r] = seed_off[r + 1] - seed_off[r];
	blockDim.x = 1; gridDim.x = 1; threadIdx.x = 0; blockIdx.x = 0;
	k_chain_node_counts(n_seeds.data(), n_reads, node_cnt.data());
	node_off[0] = 0;
	for (uint32_t r = 0; r < n_reads; ++r) node_off[r + 1] = node_off[r] + node_cnt[r];
	std::vector<uint32_t> nodes((size_t)node_off[n_reads] * 24 + 24);
	unsigned long long work = 0; int error = 0;
	ChainArgs a;
	memset(&a, 0, sizeof a);
	a.n_reads = n_reads; a.opt = *opt; a.copt = *copt; a.off = read_off; a.mem_off = mem_off; a.mems = mems; a.seed_off = seed_off;
	a.rbeg = (const uint64_t*)rbeg; a.l_pac = l_pac; a.n_seqs = n_seqs; a.c_off = c_off; a.c_alt = c_alt;
	a.s_next = s_next.data(); a.s_qb_len = s_qb_len.data(); a.chains = tmp.data(); a.order = order.data(); a.klist = klist.data();
	a.node_off = node_off.data(); a.nodes = nodes.data(); a.node_cap = node_off[n_reads]; a.seed_cap = ns; a.mems_cap = mem_off[n_reads];
	a.n_chain = n_chain.data(); a.n_cseed = n_cseed.data(); a.l_rep = l_rep.data(); a.work = &work; a.error = &error;
	k_chain_build(a);
	chain_off[0] = cseed_off[0] = 0;
	for (uint32_t r = 0; r < n_reads; ++r) { chain_off[r + 1] = chain_off[r] + n_chain[r]; cseed_off[r + 1] = cseed_off[r] + n_cseed[r]; }
	blockDim.x = 8;
	for (unsigned t = 0; t < 8; ++t) { // the eight lanes of a read, one after the other
		threadIdx.x = t;
		k_chain_emit(a, chain_off, cseed_off, chain_off[n_reads], cseed_off[n_reads], chains, s_lo, s_hi, s_qbeg, s_len);
	}
	return error;
}

}
