/* #included into the SCRATCH COPY of mapping/bwamem.c (integration/build_bwamem_gpu.sh), right before mem_align1_core:
 * mem_chain_t / mem_chain_v are private types of that file (bwamem.c:278-291), so the few lines that rebuild a read's
 * mem_chain_v from the chains the GPUs produced (CSGPU_CHAIN=1) have to live inside it.  The result is what
 * mem_chain() followed by mem_chain_flt() returns (bwamem.c:359-497): same chains, same order, same fields. */
static mem_chain_v csgpu_chains_of_read(int l_seq)
{
	mem_chain_v chain;
	int c, n = csgpu_n_chains();
	kv_init(chain);
	if (n == 0) return chain;
	kv_resize(mem_chain_t, chain, n);
	for (c = 0; c < n; ++c) {
		mem_chain_t *p = &chain.a[c];
		int rid, w, kept, is_alt, ns, l_rep, j;
		uint64_t s0;
		csgpu_chain(c, &rid, &w, &kept, &is_alt, &ns, &l_rep, &s0);
		memset(p, 0, sizeof(mem_chain_t));
		p->n = ns; p->m = ns < 4 ? 4 : ns; p->first = -1; p->rid = rid; p->w = w; p->kept = kept; p->is_alt = is_alt;
		p->frac_rep = (float)l_rep / l_seq;
		p->seeds = calloc(p->m, sizeof(mem_seed_t));
		for (j = 0; j < ns; ++j) {
			int64_t rbeg; int qbeg, len;
			csgpu_chain_seed(s0 + j, &rbeg, &qbeg, &len);
			p->seeds[j].rbeg = rbeg; p->seeds[j].qbeg = qbeg; p->seeds[j].len = p->seeds[j].score = len;
		}
		p->pos = p->seeds[0].rbeg;
	}
	chain.n = n;
	return chain;
}
