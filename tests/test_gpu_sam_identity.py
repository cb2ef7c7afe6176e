"""End-to-end drop-in check (GPU box): the reference's own bwamem host pipeline with its seeding
replaced by the compseed_b200 C-ABI (integration/_build/bwamem_gpu) must write
a SAM that is byte-identical to the unmodified `bwamem` and `CompSeed` binaries for the same -K."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from compseed_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref")
GPUBIN = os.path.join(ROOT, "integration", "_build", "bwamem_gpu")


def _have():
    return all(os.path.exists(p) for p in [GPUBIN] + [os.path.join(REFBIN, b) for b in ("bwaidx", "bwamem", "CompSeed")])


def _run(binary, args, out, env=None):
    with open(out, "wb") as f:
        subprocess.run([binary] + args, stdout=f, stderr=subprocess.DEVNULL, check=True, timeout=900,
                       env=dict(os.environ, **env) if env else None)
    return hashlib.md5(open(out, "rb").read()).hexdigest()


@pytest.mark.skipif(not _have(), reason="needs oracle/_ref binaries and integration/_build/bwamem_gpu (built where /root/reference exists)")
@pytest.mark.parametrize("kind", ["random", "repeat"])
def test_sam_is_byte_identical(cuda_lib, tmp_path, kind):
    d = str(tmp_path)
    if kind == "random":
        ref = synth.random_reference(300_000, seed=301)
        bases, off, _ = synth.simulate_reads(ref, 6000, [100, 150, 250], 0.01, seed=302, n_rate=0.001)
    else:   # repeats: x[2] > max_occ, many seeds per mem, chains across duplications
        ref = synth.repeat_rich_reference(200_000, seed=303, n_segdup=60, segdup_len=2000, n_tandem=40)
        bases, off, _ = synth.simulate_reads(ref, 3000, [100, 150], 0.02, seed=304)
    synth.write_fasta(os.path.join(d, "ref.fa"), ref)
    synth.write_reads_txt(os.path.join(d, "reads.txt"), bases, off)
    subprocess.run([os.path.join(REFBIN, "bwaidx"), "-p", os.path.join(d, "ref"), os.path.join(d, "ref.fa")],
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    idx, reads = os.path.join(d, "ref"), os.path.join(d, "reads.txt")
    want = _run(os.path.join(REFBIN, "bwamem"), ["-t", "4", "-K", "200000", idx, reads], os.path.join(d, "bwamem.sam"))
    got = _run(GPUBIN, ["-t", "4", "-K", "200000", idx, reads], os.path.join(d, "gpu.sam"))
    assert os.path.getsize(os.path.join(d, "gpu.sam")) > 100_000
    assert got == want, "GPU-seeded SAM differs from bwamem's"
    if kind == "random":   # CompSeed's own binary agrees too (it crashes on some repeat-rich inputs)
        cs = _run(os.path.join(REFBIN, "CompSeed"), ["-t", "4", "-K", "200000", idx, reads], os.path.join(d, "compseed.sam"))
        assert cs == want
    # a different -K only moves batch boundaries: same SAM for single-end input
    got2 = _run(GPUBIN, ["-t", "2", "-K", "50000", idx, reads], os.path.join(d, "gpu2.sam"))
    assert got2 == want
    # without the batch-ahead prefetch from the reader step (every batch seeded at the start of its own step 1), and with
    # pipeline batches far smaller than the -K batch (many batches per set, several sets)
    got3 = _run(GPUBIN, ["-t", "4", "-K", "200000", idx, reads], os.path.join(d, "gpu3.sam"), env={"CSGPU_NO_PREFETCH": "1"})
    assert got3 == want
    got4 = _run(GPUBIN, ["-t", "3", "-K", "30000", idx, reads], os.path.join(d, "gpu4.sam"), env={"CSGPU_BATCH": "700"})
    assert got4 == want
    # every visible GPU (the index replicated device-to-device, one contiguous block of each batch per GPU) or just one
    got5 = _run(GPUBIN, ["-t", "4", "-K", "100000", idx, reads], os.path.join(d, "gpu5.sam"), env={"CSGPU_DEVICES": "all", "CSGPU_BATCH": "1024"})
    got6 = _run(GPUBIN, ["-t", "4", "-K", "100000", idx, reads], os.path.join(d, "gpu6.sam"), env={"CSGPU_DEVICES": "1"})
    assert got5 == want and got6 == want
    # chaining and chain filtering on the GPUs too: mem_align1_core takes the chains instead of running
    # mem_chain / mem_chain_flt; only chains cross the device-to-host link
    got7 = _run(GPUBIN, ["-t", "4", "-K", "200000", idx, reads], os.path.join(d, "gpu7.sam"), env={"CSGPU_CHAIN": "1"})
    got8 = _run(GPUBIN, ["-t", "3", "-K", "40000", idx, reads], os.path.join(d, "gpu8.sam"), env={"CSGPU_CHAIN": "1", "CSGPU_BATCH": "900", "CSGPU_DEVICES": "all"})
    assert got7 == want and got8 == want
    # non-default seeding options travel through the shim
    a = _run(os.path.join(REFBIN, "bwamem"), ["-t", "4", "-k", "15", "-r", "1.0", "-y", "40", "-c", "50", idx, reads], os.path.join(d, "a.sam"))
    b = _run(GPUBIN, ["-t", "4", "-k", "15", "-r", "1.0", "-y", "40", "-c", "50", idx, reads], os.path.join(d, "b.sam"))
    assert a == b
    # ... and non-default chaining options (-w band width, -D drop ratio, -W min chain weight, -G max chain extend)
    ca = _run(os.path.join(REFBIN, "bwamem"), ["-t", "4", "-w", "40", "-D", "0.7", "-W", "25", idx, reads], os.path.join(d, "ca.sam"))
    cb = _run(GPUBIN, ["-t", "4", "-w", "40", "-D", "0.7", "-W", "25", idx, reads], os.path.join(d, "cb.sam"), env={"CSGPU_CHAIN": "1"})
    assert ca == cb
