// Types and kernels of the chaining stage (cs_chain.cu).  Deliberately free of device intrinsics and of cs_device.cuh, so
// that tests/emul/chain_emul.cpp can compile cs_chain.cu as plain C++ and run its kernels serially on the CPU against the
// reference (test infrastructure: the product never runs that build).
#pragma once
#include <stdint.h>
#include "../../include/compseed_b200.h"

struct ChainTmp {              // a chain while it is being built / filtered (per potential chain: at most one per seed)
	uint32_t first, last, n;   // first / last seed (index inside the read), number of seeds
	int32_t rid;
	uint32_t w;                // mem_chain_weight
	int32_t first_sh;          // mem_chain_t.first: the first chain this one shadows (index into the sorted order), -1 if none
	uint32_t kept;
	uint32_t pad;
};

struct ChainArgs {
	uint32_t n_reads;
	cs_seed_opt_t opt;
	cs_chain_opt_t copt;
	const uint32_t *off;         // read offsets (read length)
	const uint32_t *mem_off; const cs_mem_t *mems;      // sorted mems of the batch
	const uint32_t *seed_off; const uint64_t *rbeg;     // resolved seed positions, emission order
	int64_t l_pac; int32_t n_seqs; const int64_t *c_off; const uint8_t *c_alt;   // contigs
	uint32_t *s_next;            // [seeds] next seed of the same chain
	uint32_t *s_qb_len;          // [seeds] qbeg << 16 | len
	ChainTmp *chains;            // [seeds]
	uint32_t *order, *klist;     // [seeds] chains in traversal / sorted / output order; indices of the kept chains while filtering
	const uint32_t *node_off;    // [n_reads+1] B-tree node region of each read
	uint32_t *nodes;
	uint64_t node_cap, seed_cap, mems_cap;   // capacities of nodes[] / the per-seed arrays / the sorted mems (a batch past them is reported, not chained)
	uint32_t *n_chain, *n_cseed, *l_rep;   // [n_reads] per read: chains kept, their seeds, repetitive bases
	unsigned long long *work;
	int *error;
};

__global__ void k_chain_node_counts(const uint32_t *read_n_seeds, uint32_t n_reads, uint32_t *out);
__global__ void k_chain_build(ChainArgs a);
__global__ void k_chain_emit(ChainArgs a, const uint32_t *chain_off, const uint32_t *cseed_off, uint64_t chain_cap, uint64_t cseed_cap,
                             cs_chain_t *out, uint32_t *s_lo, uint8_t *s_hi, uint16_t *s_qbeg, uint16_t *s_len);

