#!/bin/bash
# Builds integration/_build/bwamem_gpu: the reference's bwamem with its seeding replaced by the
# compseed_b200 C-ABI.  Host code stays unchanged: the one-line edits below are applied to SCRATCH
# COPIES of mapping/bwamem.c and mapping/fastmap.c (never committed); every other file compiles straight from $REF.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"; OUT="$HERE/_build"; OBJ="$ROOT/oracle/_ref/obj"
[ -f "$REF/mapping/bwamem.c" ] || { echo "reference sources not present at $REF: keeping prebuilt $OUT"; exit 0; }
make -s -C "$ROOT/oracle" ref -j8
python "$ROOT/compseed_b200/build.py" > /dev/null
mkdir -p "$OUT"
sed -e 's|^#include "bwamem.h"|#include "bwamem.h"\n#include "cs_shim.h"|' \
    -e 's|^\tmem_collect_intv(opt, bwt, len, seq, aux, tid);|\tcsgpu_fill_mems(\&aux->mem); /* was: mem_collect_intv(opt, bwt, len, seq, aux, tid) */|' \
    -e 's|s.rbeg = tmp.pos = bwt_sa(bwt, p->x\[0\] + k);|s.rbeg = tmp.pos = csgpu_next_rbeg(); /* was: bwt_sa(bwt, p->x[0] + k) */|' \
    -e 's|^\t\tw->regs\[i\] = mem_align1_core(|\t\tcsgpu_set_read(i); w->regs[i] = mem_align1_core(|' \
    -e 's|^\t\tw->regs\[i<<1\|0\] = mem_align1_core(|\t\tcsgpu_set_read(i<<1\|0); w->regs[i<<1\|0] = mem_align1_core(|' \
    -e 's|^\t\tw->regs\[i<<1\|1\] = mem_align1_core(|\t\tcsgpu_set_read(i<<1\|1); w->regs[i<<1\|1] = mem_align1_core(|' \
    -e 's|^\tkt_for(opt->n_threads, worker1, &w, |\tcsgpu_seed_batch(opt, bwt, bns, n, seqs);\n\tkt_for(opt->n_threads, worker1, \&w, |' \
    -e 's|^mem_alnreg_v mem_align1_core(|#include "bwamem_chain_glue.c"\n\nmem_alnreg_v mem_align1_core(|' \
    -e 's|^\tchn = mem_chain(opt, bwt, bns, l_seq, (uint8_t\*)seq, buf, tid);|\tchn = csgpu_chaining() ? csgpu_chains_of_read(l_seq) : mem_chain(opt, bwt, bns, l_seq, (uint8_t*)seq, buf, tid);|' \
    -e 's|^\tchn.n = mem_chain_flt(opt, chn.n, chn.a);|\tif (!csgpu_chaining()) chn.n = mem_chain_flt(opt, chn.n, chn.a);|' \
    "$REF/mapping/bwamem.c" > "$OUT/bwamem_gpu.c"
for pat in 'cs_shim.h' 'csgpu_fill_mems' 'csgpu_next_rbeg' 'csgpu_set_read(i);' 'csgpu_seed_batch' 'bwamem_chain_glue.c' 'csgpu_chains_of_read(l_seq)' 'if (!csgpu_chaining()) chn.n'; do
  grep -q "$pat" "$OUT/bwamem_gpu.c" || { echo "patch site not found: $pat"; exit 1; }
done
diff -u "$REF/mapping/bwamem.c" "$OUT/bwamem_gpu.c" > "$OUT/bwamem_gpu.patch" || true
# batch-ahead overlap: step 0 of process() (the reader, fastmap.c:76-103) hands the batch it just read to the GPUs
sed -e 's|^#include "bwamem.h"|#include "bwamem.h"\n#include "cs_shim.h"|' \
    -e 's|^\t\treturn ret;|\t\tcsgpu_prefetch_batch(aux->opt, aux->idx->bwt, aux->idx->bns, ret->n_seqs, ret->seqs); /* added: seed batch i+1 while batch i is chained */\n\t\treturn ret;|' \
    "$REF/mapping/fastmap.c" > "$OUT/fastmap_gpu.c"
for pat in 'cs_shim.h' 'csgpu_prefetch_batch'; do
  grep -q "$pat" "$OUT/fastmap_gpu.c" || { echo "patch site not found in fastmap.c: $pat"; exit 1; }
done
diff -u "$REF/mapping/fastmap.c" "$OUT/fastmap_gpu.c" >> "$OUT/bwamem_gpu.patch" || true
CF="-O3 -g0 -fcommon -mavx2 -w -I$REF -I$REF/mapping -I$HERE -I$ROOT/include"
gcc $CF -c "$OUT/bwamem_gpu.c" -o "$OUT/bwamem_gpu.o"
gcc $CF -I"$REF/bwalib" -c "$OUT/fastmap_gpu.c" -o "$OUT/fastmap_gpu.o"
gcc $CF -c "$HERE/cs_shim.c" -o "$OUT/cs_shim.o"
gcc -o "$OUT/bwamem_gpu" "$OUT/bwamem_gpu.o" "$OUT/fastmap_gpu.o" "$OUT/cs_shim.o" \
    "$OBJ"/cstl/kstring.o "$OBJ"/cstl/kthread.o "$OBJ"/FM_index/{bntseq,bwt,bwt_gen,is,QSufSort,rle,rope}.o \
    "$OBJ"/bwalib/{bwashm,bwa,kopen,ksw,utils}.o "$OBJ"/mapping/{bwamem_pair,bwamem_extra}.o \
    -L"$ROOT/compseed_b200/_lib" -lcompseed_b200 -Wl,-rpath,'$ORIGIN/../../compseed_b200/_lib' -Wl,-rpath,/usr/local/cuda/lib64 \
    -L/usr/local/cuda/lib64 -lcudart -lstdc++ -lm -lz -lpthread -lrt
echo "built $OUT/bwamem_gpu"; grep -c '^[+-][^+-]' "$OUT/bwamem_gpu.patch" | sed 's/^/changed lines in bwamem.c: /'
