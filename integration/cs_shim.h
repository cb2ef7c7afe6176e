/* Reference-side binding of compseed_b200: the five calls a maintainer adds to mapping/bwamem.c
 * (integration/build_bwamem_gpu.sh applies them to scratch copies of bwamem.c and fastmap.c). */
#ifndef CS_SHIM_H
#define CS_SHIM_H
#include <stdint.h>
#include "FM_index/bwt.h"
#include "mapping/bwamem.h"

#ifdef __cplusplus
extern "C" {
#endif
/* fastmap.c:98, end of step 0 of process(): hand the batch just read to the GPUs and return at once (batch i+1 is seeded
 * while the host chains / extends batch i) */
void csgpu_prefetch_batch(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, int n, const bseq1_t *seqs);
/* bwamem.c:1343, before kt_for(worker1): wait for the batch's seeds (or seed it now if it was not prefetched) */
void csgpu_seed_batch(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, int n, const bseq1_t *seqs);
/* CSGPU_CHAIN=1: mem_chain + mem_chain_flt also run on the GPUs; mem_align1_core (bwamem.c:1179-1181) takes its chains from
 * integration/bwamem_chain_glue.c, which reads them through these accessors */
int csgpu_chaining(void);
int csgpu_n_chains(void);
void csgpu_chain(int c, int *rid, int *w, int *kept, int *is_alt, int *n, int *l_rep, uint64_t *first_seed);
void csgpu_chain_seed(uint64_t s, int64_t *rbeg, int *qbeg, int *len);
/* bwamem.c:1299-1305, worker1: which read the calling thread is about to align */
void csgpu_set_read(int i);
/* bwamem.c:373, replaces mem_collect_intv(opt, bwt, len, seq, aux, tid): aux->mem = sorted mems of the read */
void csgpu_fill_mems(bwtintv_v *mem);
/* bwamem.c:398, replaces bwt_sa(bwt, p->x[0] + k): next resolved position in emission order */
int64_t csgpu_next_rbeg(void);
void csgpu_destroy(void);
#ifdef __cplusplus
}
#endif
#endif
