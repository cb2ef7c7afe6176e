// Private to the library: the index handle and the error helpers shared by cs_api.cu, cs_verify.cu and cs_multi.cu.
#pragma once
#include <cstdarg>
#include "cs_kernels.cuh"

struct cs_index {
	int device;
	DevIndex d;
	uint4 *d_buckets;
	uint64_t *d_sa;
	uint4 *d_kt;            // top-of-search table (depths 1..d.kt_depth)
	uint32_t *d_pt;         // occurrence filter (2-bit counts of all d.pt_k-mers)
	uint64_t *d_text, *d_isa; // unique-match fast path: 2-bit text and sampled inverse SA
	uint8_t *d_rep;         // repeat lengths (one byte per text position)
	uint64_t bytes;
	uint64_t bwt_size_ref;  // words of the reference layout
	int sa_intv;
	int n_sm;
	cs_index_config_t cfg;
};

int cs_set_err(int code, const char *fmt, ...);
int cs_use_device(int device);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	cs_set_err(CS_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); goto fail; } } while (0)

// cs_api.cu internals used by the multi-device pipeline (cs_multi.cu)
struct cs_ctx;
int cs_i_submit(cs_ctx *ctx, int slot, uint32_t n, const uint64_t *off64, const uint8_t *bases, const uint64_t *packed, const uint32_t *nmask,
                uint64_t r0, const cs_seed_opt_t *opt);
int cs_i_finish(cs_ctx *ctx, int slot, uint64_t *n_mems, uint64_t *n_seeds);   // (chains / chain seeds when the batch was chained)
int cs_i_fetch_chains_into(cs_ctx *ctx, int slot, uint32_t *chain_off, uint32_t *cseed_off, cs_chain_t *ch, uint32_t *lo, uint8_t *hi, uint16_t *qb, uint16_t *ln);
int cs_i_fetch_compact_into(cs_ctx *ctx, int slot, uint32_t *mem_off, uint32_t *seed_off, cs_cmem_t *cm, uint32_t *lo, uint8_t *hi);
int cs_i_fetch_wait(cs_ctx *ctx, int slot, cs_counters_t *cnt, float *slot_ms);
int cs_i_poll(cs_ctx *ctx, int slot);
int cs_i_slot_times(cs_ctx *ctx, int slot, float *t5, long long *base_host_ns);
void cs_i_ctx_caps(const cs_ctx *ctx, uint64_t *max_mems, uint64_t *max_seeds);
const cs_index *cs_i_ctx_index(const cs_ctx *ctx);
