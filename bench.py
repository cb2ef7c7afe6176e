#!/usr/bin/env python
"""Benchmark of the SMEM seeding hot path (BASELINE.json metric: seeding reads/s; Occ lookups/s vs
the HBM sector peak).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU seeding on the host cores

A step = one pass of the seeding path (3 rounds + SA resolution) over one batch of synthetic reads.
At N=1 the workload is BASELINE.json configs[1]: a 3.1 Gbp i.i.d. reference (hg19-sized, 6.2 G BWT
rows) and 10 M position-sorted 150-bp reads with 1 % substitutions.  At N>1 every rank holds a
replica of the index and seeds its own contiguous block of 10 M reordered reads (weak scaling, no
collective on the data path: nothing is reduced across GPUs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# one hardware queue per stream of the slot pipeline (the default of 8 makes slots wait for each other; see cs_api.cu); must be set
# before the CUDA context exists, i.e. before torch touches the device
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-len", type=int, default=3_100_000_000, help="reference length in bp (config 2: 3.1 Gbp)")
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--sa-intv", type=int, default=1, help="device SA sampling (1 = dense; 32 = the reference's on-disk sampling)")
    ap.add_argument("--e2e-batch", type=int, default=1 << 21, help="reads per pipelined batch on the host-buffer path")
    ap.add_argument("--e2e-slots", type=int, default=3, help="pipelined batches in flight on the host-buffer path")
    ap.add_argument("--cpu-sample", type=int, default=300_000, help="reads of the same workload timed on the host cores")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-probe", action="store_true", help="skip the in-run random 32-byte sector gather peak (the roofline denominator)")
    ap.add_argument("--verify-stride", type=int, default=1, help="cs_index_verify on the index the run uses: every n-th row (1 = all rows, 0 = skip)")
    ap.add_argument("--e2e-input", default="packed", choices=["packed", "bytes"],
                    help="what crosses the host-to-device link on the host-buffer path: 2-bit packed reads (cs_seed_batch_submit_packed) or nt4 bytes")
    ap.add_argument("--l2-persist-mb", type=int, default=0, help="cs_ctx_config_t.l2_persist_mb for the contexts of this run")
    ap.add_argument("--overlap", action="store_true", help="cs_ctx_config_t.overlap_streams = 1")
    ap.add_argument("--e2e-only", action="store_true", help="experiments: only the host-buffer leg (prints its dict, not a bench line)")
    ap.add_argument("--isa-intv", type=int, default=-1, help="cs_index_config_t.isa_intv (sampling of the inverse SA; -1 = default 2)")
    ap.add_argument("--lit-ctas", type=int, default=-1, help="cs_ctx_config_t.lit_ctas_per_sm")
    ap.add_argument("--batch-order", type=int, default=-1, help="cs_ctx_config_t.batch_order")
    ap.add_argument("--no-bsw", action="store_true", help="skip the extension-stage leg (SURVEY 8f-2: banded Smith-Waterman, reported next to the headline)")
    ap.add_argument("--bsw-pairs", type=int, default=2_000_000, help="sequence pairs of the extension-stage leg")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.p = gpu, [], None
        self.t0 = self.t1 = None   # the timed region (host clock); only samples that arrived inside it are used

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "40"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.p.terminate()
        inside = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.05)]
        rows = inside if inside else [r for _, r in self.rows[-3:]]   # a region shorter than the polling period: the nearest samples
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "samples_inside_timed_region": len(inside), "polling_ms": 40}


def make_workload(args, rank: int, world: int, device: str):
    """Reference (same on every rank) and this rank's contiguous block of reordered reads."""
    import torch
    from compseed_b200 import synth
    ref_t = synth.random_reference_torch(args.ref_len, 20261018, device)
    lo = args.ref_len * rank // world
    hi = args.ref_len * (rank + 1) // world
    bases, off, _ = synth.simulate_reads_torch(ref_t, args.reads, args.read_len, 0.01, seed=1000 + rank, window=(lo, hi))
    ref = ref_t.cpu().numpy()
    del ref_t
    torch.cuda.empty_cache()
    return ref, bases, off


def workload_name(args, world: int) -> str:
    return (f"synthetic {args.ref_len / 1e9:.2f} Gbp i.i.d. reference (seq_len {2 * args.ref_len}), "
            f"{args.reads} position-sorted {args.read_len}bp reads/GPU x {world} GPU, 1% subst, -k19 -r1.5 -y20 -c500")


def cpu_seed_sample(host_idx, bases, off, n_sample: int, threads: int):
    """Times the reference's own CPU seeding (oracle/_ref, CompSeed SST path and the bwamem path)
    or, when oracle/_ref is absent, our C port, on the first n_sample reads of the workload.
    Returns (info, {name: result}, n): the results are what the GPU output is compared with (parity)."""
    from oracle import oracle_py as O
    n = min(n_sample, off.shape[0] - 1)
    b, o = bases[:int(off[n])], off[:n + 1]
    out = {"cores": threads, "sample": f"first {n} reads of the workload, {threads} host threads, same index"}
    if O.have_ref():
        ri = O.RefIndex.from_arrays(host_idx["primary"], host_idx["L2"], host_idx["seq_len"], host_idx["bwt"], host_idx["sa"], host_idx["sa_intv"])
        cs = ri.seed(b, o, "compseed", n_threads=threads)
        bw = ri.seed(b, o, "bwamem", n_threads=threads)
        out.update(kind="reference", value=n / cs.seconds, unit="reads/s", bwamem_reads_per_s=n / bw.seconds,
                   compseed_counters=cs.counters, seconds=cs.seconds)
        res = {"CompSeed collect_mem_with_sst/tem_forward_sst + SAL (comp_seed.cpp:67,141,2255-2346)": cs,
               "bwamem bwt_smem1/bwt_seed_strategy1/bwt_sa (bwt.c:289-379,86-96)": bw}
    else:
        oi = O.OracleIndex.from_arrays(host_idx["primary"], host_idx["L2"], host_idx["seq_len"], host_idx["bwt"], host_idx["sa"], host_idx["sa_intv"])
        r = oi.seed(b, o, n_threads=threads)
        out.update(kind="port", value=n / r.seconds, unit="reads/s", seconds=r.seconds)
        res = {"oracle port (oracle/cs_oracle.c)": r}
    return out, res, n


def contig_lens(ref_len: int):
    """The reference sequences the chaining stage sees: bntann1_t.len is int32 (bntseq.h:43), so the 3.1 Gbp reference is a
    set of 25 equal contigs, human-chromosome sized; seeding never looks at contig boundaries."""
    n = max(1, -(-ref_len // 124_000_000))
    base = ref_len // n
    return [base] * (n - 1) + [ref_len - base * (n - 1)]


def same_prefix(got, want, n: int) -> bool:
    """The first n reads of `got` (mem_off, mems, seed_off, rbeg of a longer run) == `want` (exactly n reads), bit for bit."""
    nm, ns = int(want.mem_off[-1]), int(want.seed_off[-1])
    if got.mem_off.shape[0] < n + 1 or int(got.mem_off[n]) != nm or int(got.seed_off[n]) != ns:
        return False
    return bool(np.array_equal(got.mem_off[:n + 1], want.mem_off) and np.array_equal(got.seed_off[:n + 1], want.seed_off)
                and np.array_equal(got.mems[:nm], want.mems) and np.array_equal(got.rbeg[:ns], want.rbeg))


def work_counters(host_idx, bases, off, n_sample: int, threads: int):
    """Implementation-independent work per read from the instrumented C port."""
    from oracle import oracle_py as O
    n = min(n_sample, off.shape[0] - 1)
    oi = O.OracleIndex.from_arrays(host_idx["primary"], host_idx["L2"], host_idx["seq_len"], host_idx["bwt"], host_idx["sa"], host_idx["sa_intv"])
    r = oi.seed(bases[:int(off[n])], off[:n + 1], n_threads=threads)
    c = r.counters
    return {k: c[k] / n for k in c}, r, n


def e2e_only(args, cs, idx, bases, off, opt, ccfg, threads):
    """Experiments on the host-buffer path: one line per (input form) with where the pipeline's host thread spent its time."""
    n_reads = off.shape[0] - 1
    off64 = off.astype(np.uint64)
    for packed_in in ((True, False) if args.e2e_input == "packed" else (False,)):
        if packed_in:
            pk, nm = cs.pack_reads_host64(bases, off64, threads)
            cs.host_register(pk); cs.host_register(nm)
        else:
            cs.host_register(bases)
        ms_ = cs.MultiSeeder([idx], batch_reads=min(args.e2e_batch, n_reads), max_read_len=args.read_len, n_slots=args.e2e_slots,
                             mems_per_read=14, seeds_per_read=20, config=ccfg)
        sub = (lambda s: ms_.submit_packed(s, pk, nm, off64, opt)) if packed_in else (lambda s: ms_.submit(s, bases, off64, opt))
        for _ in range(2):
            sub(0); sub(1); ms_.wait(0, gather=False); ms_.wait(1, gather=False)
        t0 = time.perf_counter()
        sub(0)
        infos = []
        for i in range(1, args.steps):
            sub(i & 1)
            infos.append(ms_.wait((i - 1) & 1, gather=False))
        infos.append(ms_.wait((args.steps - 1) & 1, gather=False))
        dt = time.perf_counter() - t0
        tr = ms_.trace((args.steps - 1) & 1)   # timeline of the last set: ms relative to its first submit
        tr0 = float(tr[:, 5].min()) if tr.size else 0.0
        hs = {k: float(np.mean([x["host_s"][k] for x in infos])) for k in infos[0]["host_s"]}
        hs.update({"gpu_ms_" + k: float(np.mean([x["gpu_ms"][k] for x in infos])) for k in infos[0]["gpu_ms"]})
        print(json.dumps({"e2e_reads_per_s": n_reads * args.steps / dt, "input": "packed" if packed_in else "bytes", "batch": args.e2e_batch, "slots": args.e2e_slots,
                          "set_seconds": float(np.mean([x["seconds"] for x in infos])), "host_thread_s_per_set": hs, "wire_bytes": infos[-1]["wire_bytes"],
                          "timeline_cols": "gpu: h2d_enqueued kernels_start kernels_end d2h_start d2h_end | host: submit kernels_seen results_seen",
                          "timeline_ms": [[round(float(v) - tr0, 2) for v in row] for row in tr]}))
        ms_.close()
        if packed_in:
            cs.host_unregister(pk); cs.host_unregister(nm)
        else:
            cs.host_unregister(bases)
    return 0


def bsw_leg(args, cs, device: int, threads: int):
    """SURVEY 8f-2, measured to the same bar as the seeding path: cs_bsw_* (one ksw_extend2 per pair) on a synthetic batch of extension
    pairs -- device-resident cells/s from the kernel's CUDA events, pairs/s through the C-ABI with host buffers, and the unmodified
    reference's own batch routines on the host cores over a bounded sample of the same pairs, whose results are compared bit for bit."""
    from compseed_b200 import synth
    from oracle import oracle_py as O
    from concurrent.futures import ThreadPoolExecutor
    pairs, ref, qer = synth.extension_pairs_fast(args.bsw_pairs, seed=411)
    n = pairs.shape[0]
    ex = cs.BswExtender(device, n, ref.nbytes, qer.nbytes, 256)
    ex.stage(pairs, ref, qer)
    l0 = ex.launches
    ms_best, cells = None, 0
    for it in range(args.warmup + args.steps):
        ms, cells = ex.run_staged()
        if it >= args.warmup:
            ms_best = ms if ms_best is None or ms < ms_best else ms_best
    launches = ex.launches - l0
    got = pairs.copy()
    t0 = time.perf_counter()
    ex.extend(got, ref, qer)
    e2e_s = time.perf_counter() - t0
    ex.close()
    out = {"what": "banded Smith-Waterman extension (ksw_extend2, bwalib/ksw.c:380 == BandedPairWiseSW, mapping/bandedSWA.cpp), one pair per thread, 16-bit cells",
           "pairs": n, "cells": int(cells), "kernel_ms": ms_best, "gcups": cells / ms_best / 1e6, "pairs_per_s": n / (ms_best * 1e-3),
           "e2e_pairs_per_s": n / e2e_s, "e2e_what": "cs_bsw_extend with host buffers (SeqPair array + two sequence buffers in, scores out), wall clock",
           "h2d_bytes": int(pairs.nbytes + ref.nbytes + qer.nbytes), "d2h_bytes": int(pairs.nbytes), "gpu_launches": int(launches),
           "bound": "integer ALU / latency: one dependent chain of rows per thread; DP rows live in an interleaved HBM scratch that stays in L2 "
                    "(profiles/r02_ncu_k_bsw_*.txt); not tensor-core work (max-plus recurrences)"}
    if O.have_ref():   # the reference's own routines on the host cores, parallel chunks of a bounded sample of the same pairs
        ns = min(n, 200_000)
        sub = np.ascontiguousarray(pairs[:ns])

        def run_ref(jobs):   # jobs: (mode, indices); returns (results in place of the pairs, seconds)
            res_all = sub.copy()
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as tp:
                res = list(tp.map(lambda j: O.ref_bsw(np.ascontiguousarray(sub[j[1]]), ref, qer, mode=j[0])[0], jobs))
            dt = time.perf_counter() - t0
            for (mode, c), r in zip(jobs, res):
                res_all[c] = r
            return res_all, dt

        # (a) the parity target: scalarBandedSWA == ksw_extend2, the routine bwamem calls and the one the SIMD twins are written after
        want0, sec0 = run_ref([(0, c) for c in np.array_split(np.arange(ns), threads) if c.size])
        # (b) what CompSeed times: getScores8 on the pairs that fit 8-bit cells, getScores16 on the rest (comp_seed.cpp:1556-1564)
        m8 = (sub[:, 3] < 128) & (sub[:, 4] < 128) & (sub[:, 5] + np.minimum(sub[:, 3], sub[:, 4]) < 128)
        jobs = []
        for mode, idx in ((1, np.flatnonzero(m8)), (2, np.flatnonzero(~m8))):
            share = max(1, int(round(threads * idx.size / ns)))
            jobs += [(mode, c) for c in np.array_split(idx, share) if c.size]
        want_simd, cpu_s = run_ref(jobs)
        cells_s = O.oracle_bsw(sub, ref, qer, n_threads=threads)[1]
        out["cpu_baseline"] = {"kind": "reference", "cores": threads, "value": ns / cpu_s, "unit": "pairs/s", "gcups": cells_s / cpu_s / 1e9,
                               "sample": "first %d pairs, BandedPairWiseSW::getScores8 / getScores16 (mapping/bandedSWA.cpp, AVX2 build) by size class, %d parallel chunks" % (ns, len(jobs)),
                               "scalar_routine_pairs_per_s": ns / sec0}
        simd_diff = int((want_simd[:, 8:] != want0[:, 8:]).any(axis=1).sum())
        out["parity"] = {"pairs": ns, "against": "BandedPairWiseSW::scalarBandedSWAWrapper (== ksw_extend2, bwalib/ksw.c:380): score, tle, gtle, qle, gscore, max_off of every pair",
                         "equal": bool(np.array_equal(got[:ns][:, 8:], want0[:, 8:])),
                         "reference_simd_twins_differ_from_its_scalar_routine_on_pairs": simd_diff,
                         "simd_note": "the reference's getScores8/16 are not bit-identical to its own scalar routine: on a few pairs whose extension ends at once "
                                      "their gtle / gscore differ (and depend on which pairs share a SIMD batch); score, tle, qle and max_off agree",
                         "equal_to_simd_twins_on_score_tle_qle_max_off": bool(np.array_equal(got[:ns][:, [8, 9, 11, 13]], want_simd[:, [8, 9, 11, 13]]))}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1

    if args.impl == "reference" and rank != 0:
        return 0  # rank 0 alone runs the CPU arm

    import torch
    import compseed_b200 as cs
    from compseed_b200 import build as B
    B.build()
    if not torch.cuda.is_available() or cs.device_count() == 0:
        print(json.dumps({"error": "no CUDA device: compseed_b200 has no CPU path"}))
        return 2
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))

    t_setup = time.time()
    ref, bases, off = make_workload(args, rank if args.impl == "ours" else 0, world if args.impl == "ours" else 1, device)
    n_reads = off.shape[0] - 1
    idx = cs.FMIndex.build(ref, device=local_rank, sa_intv=args.sa_intv, config=cs.IndexConfig(isa_intv=args.isa_intv))
    opt = cs.SeedOpt()
    setup_s = time.time() - t_setup

    # ---------------------------------------------------------------------------------------
    if args.impl == "reference":
        del ref
        host_idx = idx.download(sa_intv=32)
        idx.close()
        n_s = min(args.cpu_sample, n_reads)
        vals, last = [], None
        for it in range(args.warmup + args.steps):
            info, _, n_s = cpu_seed_sample(host_idx, bases, off, n_s, threads)
            if it >= args.warmup:
                vals.append(info["seconds"])
            last = info
            if it == 0 and info["seconds"] * (args.warmup + args.steps) > 240:   # keep the whole run within minutes
                n_s = max(10_000, int(n_s * 240 / (info["seconds"] * (args.warmup + args.steps))))
        sec = float(np.mean(vals))
        v = n_s / sec
        line = {"impl": "reference", "metric": "smem_seeding_reads_per_s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(args, 1), "inputs": "host memory (CPU arm)"},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": last["kind"],
                                 "sample": f"each step: first {n_s} reads of the workload, {threads} host threads, CompSeed SST seeding + SAL",
                                 "bwamem_reads_per_s": last.get("bwamem_reads_per_s")},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "index_built_by": "cs_index_build on the GPU (bit-identical to bwaidx; the CPU builder needs hours for 3.1 Gbp)",
                "setup_s": setup_s}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------
    # our arm.  (0) the index this run rests on, checked against the definitions of its parts (the 3.1 Gbp index is
    # built on the GPU; bwaidx would need hours), and the random-sector peak of its own arrays: the roofline denominator
    index_verify = None
    if args.verify_stride > 0 and args.sa_intv == 1:
        t_v = time.time()
        index_verify = idx.verify(ref, stride=args.verify_stride)
        index_verify["seconds"] = time.time() - t_v
        index_verify["stride"] = args.verify_stride
    del ref
    probe = None
    if not args.no_probe:
        p1 = idx.probe_gather(1 << 28, 2, 1)
        p4 = idx.probe_gather(1 << 28, 2, 4)
        t1 = cs.probe_random_gather(local_rank, 16 << 30, 32, 1 << 28, 2, 1)
        t4 = cs.probe_random_gather(local_rank, 16 << 30, 32, 1 << 28, 2, 4)
        best = max(p1, p4, t1, t4, key=lambda x: x[1])
        probe = {"gloads_per_s": best[1], "gb_per_s": best[0],
                 "over_the_index_arrays": {"one_load_in_flight_per_thread_gloads_per_s": p1[1], "four_loads_in_flight_per_thread_gloads_per_s": p4[1],
                                           "bytes": idx.device_bytes},
                 "over_one_16_GiB_table": {"one_load_in_flight_per_thread_gloads_per_s": t1[1], "four_loads_in_flight_per_thread_gloads_per_s": t4[1]},
                 "what": "independent uniformly random 32-byte sector loads, 1184 CTAs x 256 threads, measured in this run before the timed region: over this "
                         "index's own arrays (cs_probe_index_gather, each load picks an array in proportion to its size) and over one 16 GiB table "
                         "(cs_probe_random_gather); the roofline peak is the best of the four"}
    ccfg = cs.CtxConfig(l2_persist_mb=args.l2_persist_mb, overlap_streams=1 if args.overlap else 0, lit_ctas_per_sm=args.lit_ctas, batch_order=args.batch_order)

    if args.e2e_only:
        return e2e_only(args, cs, idx, bases, off, opt, ccfg, threads)

    # (1) device-resident: all reads of the step staged in HBM once
    max_mems, max_seeds = n_reads * 14, n_reads * 20
    ctx = cs.SeedContext(idx, n_reads, int(off[-1]), args.read_len, max_mems, max_seeds, 1, ccfg)
    ctx.stage(0, bases, off)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(args.warmup):
        ctx.run_staged(0, opt)
        last = ctx.wait_device(0)
    barrier()
    launches0 = ctx.launches
    sampler.mark_begin()
    t0 = time.perf_counter()
    ms = {k: 0.0 for k in ("dev", "seed_span", "pack", "fast", "walk", "lit", "r3", "r3_tail", "collect", "sa")}
    req = np.zeros(6, dtype=np.float64)
    for _ in range(args.steps):
        ctx.run_staged(0, opt)
        last = ctx.wait_device(0)
        km, k2 = last.kernel_ms, last.kernel_ms2
        ms["seed_span"] += km[0]; ms["collect"] += km[1]; ms["sa"] += km[2]
        ms["pack"] += k2[2]; ms["fast"] += km[6] - k2[2]; ms["walk"] += km[7]; ms["lit"] += k2[0]; ms["r3"] += k2[1]; ms["r3_tail"] += km[5]
        ms["dev"] += km[0] + km[1] + km[2]   # CUDA events on the launching stream: pack .. SA resolution
        req += np.array(last.gather_requests, dtype=np.float64)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.mark_end()
    clocks = sampler.stop()
    dev_launches = ctx.launches - launches0
    counters = last.counters
    n_mems, n_seeds = last.n_mems_device, last.n_seeds_device
    # size-independent properties on the whole step ...
    got = ctx.fetch(0, copy=False)
    starts = (got.mems[:, 3] >> np.uint64(32)).astype(np.int64)
    ends = (got.mems[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
    assert got.mem_off[-1] == n_mems and got.seed_off[-1] == n_seeds
    assert np.all(ends - starts >= opt.min_seed_len) and np.all(ends <= args.read_len) and np.all(got.mems[:, 2] >= 1)
    assert np.all(got.rbeg >= 0) and np.all(got.rbeg < idx.seq_len)
    info = got.mems[:, 3]
    same_read = np.ones(info.shape[0], dtype=bool)
    same_read[got.mem_off[1:-1][got.mem_off[1:-1] < info.shape[0]]] = False
    assert np.all((info[1:] >= info[:-1]) | ~same_read[1:]), "mems not sorted by info inside a read"
    del starts, ends, info, same_read
    # ... and the first cpu_sample reads kept for the bit-exact comparison with the reference below
    n_par = min(args.cpu_sample, n_reads)
    nm_p, ns_p = int(got.mem_off[n_par]), int(got.seed_off[n_par])
    dev_head = SimpleNamespace(mem_off=got.mem_off[:n_par + 1].copy(), mems=got.mems[:nm_p].copy(),
                               seed_off=got.seed_off[:n_par + 1].copy(), rbeg=got.rbeg[:ns_p].copy())
    del got
    ctx.close()

    dev_ms = ms["dev"]
    if use_dist:
        t = torch.tensor([dev_ms, wall_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    else:
        dev_ms_max, wall_ms_max = dev_ms, wall_ms
    total_reads = n_reads * world * args.steps
    value = total_reads / (dev_ms_max * 1e-3)

    # (2) end to end through the C-ABI with HOST buffers, H2D + D2H inside the timed region: the multi-device pipeline of the
    # library (cs_multi_*: one host thread per GPU submits batches through the slots of a ctx, results arrive by DMA in the
    # compact wire format in page-locked arrays).  Consecutive steps alternate between the two read sets the pipeline keeps
    # in flight, as a host does with batch i+1 and batch i.
    def run_e2e(packed_in: bool, keep_head: bool, chains: bool = False):
        bs = min(args.e2e_batch, n_reads)
        off64 = off.astype(np.uint64)
        if packed_in:   # the host holds the reads 2-bit packed in page-locked memory (what a packing reader would leave)
            pk, nm = cs.pack_reads_host64(bases, off64, threads)
            cs.host_register(pk); cs.host_register(nm)
        else:
            cs.host_register(bases)   # the reads sit in page-locked host memory, as the bench contract asks
        # (with chaining two batches in flight: its per-seed scratch is 3 GB per slot, and next to the 155 GB index three slots do not fit)
        n_slots = min(args.e2e_slots, 2) if chains else args.e2e_slots
        torch.cuda.empty_cache()
        ms_ = cs.MultiSeeder([idx], batch_reads=bs, max_read_len=args.read_len, n_slots=n_slots, mems_per_read=14, seeds_per_read=20, config=ccfg)
        if chains:      # mem_chain + mem_chain_flt on the GPU too: only the filtered chains come back
            ms_.set_chaining(contig_lens(args.ref_len))

        def submit(set_id):
            if packed_in:
                ms_.submit_packed(set_id, pk, nm, off64, opt)
            else:
                ms_.submit(set_id, bases, off64, opt)

        n_b = (n_reads + bs - 1) // bs
        h2d = (pk.nbytes + nm.nbytes if packed_in else bases.nbytes) + 4 * (n_reads + n_b)
        for _ in range(max(1, args.warmup - 1)):
            submit(0); submit(1)
            ms_.wait(0, gather=False); ms_.wait(1, gather=False)
        barrier()
        l0 = ms_.launches
        t0 = time.perf_counter()
        submit(0)
        for i in range(1, args.steps):
            submit(i & 1)
            info = ms_.wait((i - 1) & 1, gather=False)
        info = ms_.wait((args.steps - 1) & 1, gather=False)
        barrier()
        e_ms = (time.perf_counter() - t0) * 1e3
        n_launch = ms_.launches - l0
        assert chains or (info["n_mems"] == n_mems and info["n_seeds"] == n_seeds), "host-buffer path disagrees with the device-resident path"
        head = None
        if keep_head and chains:   # untimed: the chains of the first reads, flat, for the comparison with the reference's mem_chain_flt output
            submit(0)
            g = ms_.wait(0, gather=True)
            nc_h, nk_h = int(g.chain_off[n_par]), int(g.cseed_off[n_par])
            head = dict(chain_off=g.chain_off[:n_par + 1], rid=g.rid[:nc_h], w=g.w[:nc_h], kept=g.kept[:nc_h], n=g.n[:nc_h], l_rep=g.l_rep[:nc_h],
                        s_rbeg=g.s_rbeg[:nk_h], s_qbeg=g.s_qbeg[:nk_h], s_len=g.s_len[:nk_h])
            del g
        elif keep_head:   # untimed: one more pass, expanded to flat arrays, for the comparison with the reference
            submit(0)
            g = ms_.wait(0, gather=True, n_threads=threads)
            nm_h, ns_h = int(g.mem_off[n_par]), int(g.seed_off[n_par])
            head = SimpleNamespace(mem_off=g.mem_off[:n_par + 1].astype(np.uint32), mems=g.mems[:nm_h].copy(),
                                   seed_off=g.seed_off[:n_par + 1].astype(np.uint32), rbeg=g.rbeg[:ns_h].copy())
            del g
        if use_dist:
            t = torch.tensor([e_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t[0])
        out = {"value": total_reads / (e_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": info["wire_bytes"] * world,
               "batch_reads": bs, "slots": n_slots,
               "api": "cs_multi_submit%s / cs_multi_wait: the library's own pipeline (one host thread per GPU, no Python in the loop), two read sets in flight" % ("_packed" if packed_in else ""),
               "input": "2-bit packed reads + N mask in page-locked host memory, packed by the host outside the timed region (cs_pack_reads_host64)" if packed_in else "nt4 bytes in page-locked host memory",
               "output": ("filtered chains (mem_chain + mem_chain_flt run on the GPU: 16 B per chain, 9 B per chain seed, 8 B of offsets per read) by DMA into "
                          "page-locked host arrays (cs_multi_read_chains)") if chains else
                         ("compact wire format (20 B per mem, 5 B per seed position, 8 B of offsets per read) by DMA into page-locked host arrays; "
                          "a consumer expands a read where it uses it (cs_cmem_unpack / cs_multi_read)"),
               "timing": "host wall clock between device syncs"}
        ms_.close()
        if packed_in:
            cs.host_unregister(pk); cs.host_unregister(nm)
        else:
            cs.host_unregister(bases)
        return out, head, n_launch

    e2e, e2e_head, e2e_launches, chain_head = None, None, 0, None
    if not args.no_e2e:
        e2e, e2e_head, e2e_launches = run_e2e(args.e2e_input == "packed", True)
        # the same path with chaining + chain filtering on the GPU: what a host that takes chains would see
        try:
            ch, chain_head, _ = run_e2e(args.e2e_input == "packed", True, chains=True)
            e2e["with_chaining_on_the_gpu"] = {k: ch[k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "slots", "output")}
        except cs.CompSeedError as e:   # (the headline does not depend on this leg; every rank must still reach the collectives below)
            e2e["with_chaining_on_the_gpu"] = {"error": str(e)}
        if args.e2e_input == "packed":   # the same with nt4 bytes crossing the link, for comparison (not the headline)
            alt, _, _ = run_e2e(False, False)
            e2e["nt4_bytes_input_variant"] = {k: alt[k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "input")}

    # (2b) what the host side of the box can move: every rank copies 1 GiB host-to-device and 1 GiB device-to-host at the
    # same time (page-locked memory, two streams), all ranks together; the sum is the ceiling of the host-buffer path at
    # this N (the GPUs of a box share the host's memory and PCIe fabric)
    host_link = None
    if not args.no_e2e:
        nb = 1 << 30
        hsrc = torch.empty(nb, dtype=torch.uint8).pin_memory(); hdst = torch.empty(nb, dtype=torch.uint8).pin_memory()
        dsrc = torch.empty(nb, dtype=torch.uint8, device=device); ddst = torch.empty(nb, dtype=torch.uint8, device=device)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        best = None
        for it in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s1):
                ddst.copy_(hsrc, non_blocking=True)
            with torch.cuda.stream(s2):
                hdst.copy_(dsrc, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if use_dist:
                t = torch.tensor([dt], device=device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t[0])
            best = dt if best is None or dt < best else best
        host_link = {"aggregate_h2d_plus_d2h_gb_per_s": 2 * nb * world / best / 1e9, "per_gpu_gb_per_s": 2 * nb / best / 1e9,
                     "what": "%d rank(s), each 1 GiB host-to-device and 1 GiB device-to-host at once from page-locked memory, slowest rank, best of 3" % world}
        if e2e is not None:
            bytes_per_read = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) / (n_reads * world)
            host_link["e2e_bytes_per_read"] = bytes_per_read
            host_link["e2e_gb_per_s"] = e2e["value"] * bytes_per_read / 1e9
            host_link["reads_per_s_this_link_allows"] = host_link["aggregate_h2d_plus_d2h_gb_per_s"] * 1e9 / bytes_per_read
        del hsrc, hdst, dsrc, ddst

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return 0

    # (3) same-run CPU baseline (rank 0, N=1 only) -- and PARITY: the reference's own mems and seed positions for the
    # first cpu_sample reads of this very workload and index, against what the GPU returned above
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    stream_peak, stream_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    cpu_baseline, per_read, parity = None, None, None
    if not args.no_cpu and world == 1:
        host_idx = idx.download(sa_intv=32)
        cpu_baseline, ref_res, n_cmp = cpu_seed_sample(host_idx, bases, off, n_par, threads)
        per_read, _, n_cnt = work_counters(host_idx, bases, off, min(args.cpu_sample, 100_000), threads)
        del host_idx
        first = next(iter(ref_res.values()))
        parity = {"reads": n_cmp, "mems": int(first.mem_off[-1]), "seeds": int(first.seed_off[-1]), "against": list(ref_res.keys()),
                  "what": "mem_off, mems (x0, x1, x2, info), seed_off, rbeg of the first reads of the step, bit for bit",
                  "device_resident_equal": all(same_prefix(dev_head, w, n_cmp) for w in ref_res.values())}
        if e2e_head is not None:
            parity["e2e_equal"] = all(same_prefix(e2e_head, w, n_cmp) for w in ref_res.values())
        if chain_head is not None and cpu_baseline["kind"] == "reference":   # the reference's own mem_chain + mem_chain_flt on its own seeds
            from oracle import oracle_py as O
            want_c = O.ref_chain(off[:n_cmp + 1], first, contig_lens(args.ref_len))
            ce = bool(np.array_equal(chain_head["chain_off"], want_c.chain_off) and np.array_equal(chain_head["rid"], want_c.rid)
                      and np.array_equal(chain_head["w"], want_c.w) and np.array_equal(chain_head["kept"], want_c.kept) and np.array_equal(chain_head["n"], want_c.n)
                      and np.array_equal(chain_head["s_rbeg"], want_c.s_rbeg) and np.array_equal(chain_head["s_qbeg"], want_c.s_qbeg)
                      and np.array_equal(chain_head["s_len"], want_c.s_len))
            parity["chains"] = {"reads": n_cmp, "chains": int(want_c.chain_off[-1]), "chain_seeds": int(want_c.s_rbeg.shape[0]), "equal": ce,
                                "against": "mem_chain + mem_chain_flt of the reference (comp_seed.cpp:241-354) on the reference's own seeds"}
        parity["equal"] = bool(parity["device_resident_equal"] and parity.get("e2e_equal", True) and parity.get("chains", {}).get("equal", True))
        parity["rows_at_or_above_2^32_in_sample"] = int((first.mems[:, 0] >= np.uint64(1 << 32)).sum())

    # (4) roofline of the dominant kernel, k_seed_fast, from what it EXECUTES: memory requests counted in the kernel
    # (one per lane and load that leaves the SM: Occ sectors, filter words, table entries, SA / inverse-SA / text words),
    # 32 bytes (one sector) each, over its CUDA-event duration; the peak is the random-sector rate of the same arrays
    # measured in this run.  The reference's logical work is reported next to it, not as a fraction of peak.
    steps = args.steps
    fast_s = ms["fast"] / steps * 1e-3
    req_fast = req[0] / steps
    achieved = req_fast * 32 / fast_s / 1e9
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_k_seed_fast_traffic.json")))
        traffic = tj["dram_bytes_per_read"] * n_reads
        traffic_src = "ncu --set full capture of k_seed_fast committed under profiles/ (%s): dram__bytes_read + write per read x reads of a launch; stale if the kernel changed since" % tj.get("capture", "?")
    except Exception:
        pass
    peak = probe["gb_per_s"] if probe else None
    roofline = {"kernel": "k_seed_fast", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if peak else None,
                "peak_source": "best random 32-byte sector gather rate measured in this run (see random_sector_peak)" if probe else "not measured (--no-probe)",
                "achieved_is": "executed memory requests of the kernel (in-kernel counter) x 32 B / CUDA-event duration",
                "requests_per_read": req_fast / n_reads, "requests_per_s": req_fast / fast_s, "ms_per_launch": ms["fast"] / steps,
                "random_sector_peak": probe,
                "traffic": traffic, "traffic_source": traffic_src,
                "streaming_view": {"peak_gb_per_s": stream_peak, "peak_source": stream_src,
                                   "dram_gb_per_s": (traffic / fast_s / 1e9) if traffic else None,
                                   "frac": (traffic / fast_s / 1e9 / stream_peak) if traffic else None},
                "all_kernels": {k: {"ms_per_step": ms[n] / steps, "requests_per_read": req[i] / steps / n_reads if i is not None else None,
                                    "grequests_per_s": (req[i] / ms[n] / 1e6) if (i is not None and ms[n] > 0) else None}
                                for k, n, i in (("k_pack_reads", "pack", None), ("k_seed_fast", "fast", 0), ("k_seed_walk", "walk", 1), ("k_seed", "lit", 2),
                                                ("third pass", "r3", 3), ("collect", "collect", None), ("k_sa_resolve", "sa", 4))},
                "kernel_share_of_step": {"k_pack_reads": ms["pack"] / dev_ms, "k_seed_fast": ms["fast"] / dev_ms, "k_seed_walk": ms["walk"] / dev_ms, "k_seed": ms["lit"] / dev_ms,
                                         "third_pass_not_hidden": ms["r3_tail"] / dev_ms, "collect": ms["collect"] / dev_ms, "k_sa_resolve": ms["sa"] / dev_ms},
                "deferred_calls_per_read": counters.get("deferred_calls", 0) / n_reads}
    # the reference's logical work on the same reads (E extends, E2 of them over two buckets, S LF steps, A SA lookups, M mems)
    ref_work = None
    if per_read is not None:
        E, e2_ratio, S, A, M = per_read["ext"], per_read["ext2"] / per_read["ext"], per_read["lf"], per_read["sa"], per_read["mem"]
        seed_bytes = 64.0 * E * (1.0 + e2_ratio) + args.read_len + 32.0 * M
        path_bytes = seed_bytes + 64.0 * S + 16.0 * A
        occ_logical = 2.0 * E + S
        ref_work = {"what": "SURVEY 8d: bytes the REFERENCE's algorithm would move for these reads (64-byte buckets, sa_intv 32); the result-neutral "
                            "structures remove most of it, so this is a speed-up in work-equivalents, not a fraction of any peak",
                    "extends_per_read": E, "two_bucket_ratio": e2_ratio, "lf_steps_per_read": S, "seeding_bytes_per_read": seed_bytes,
                    "whole_path_bytes_per_read": path_bytes, "work_equivalent_gb_per_s": path_bytes * value / 1e9,
                    "occ_lookups_per_read_logical": occ_logical, "occ_lookups_per_s_logical": occ_logical * value}
    extension = None
    if not args.no_bsw and world == 1 and not args.no_cpu:
        try:
            extension = bsw_leg(args, cs, local_rank, threads)
        except Exception as e:   # the headline does not depend on this leg
            extension = {"error": repr(e)}
    occ_exec_per_read = (2.0 * counters["ext_calls"] + counters["sal_calls"]) / n_reads
    line = {"metric": "smem_seeding_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "sa_intv_device": args.sa_intv, "index_bytes_per_gpu": idx.device_bytes,
                       "result_neutral_structures": "dense SA, top-of-search k-mer table (depth <= 13), 2-bit occurrence filter (K <= 19), "
                                                    "2-bit text + sampled inverse SA for unique matches, repeat lengths (one byte per text position) (DESIGN.md section 5)",
                       "l2_policy": "inputs larger than L2 (index %.1f GB, reads %.1f GB per step)" % (idx.device_bytes / 1e9, bases.nbytes / 1e9),
                       "l2_persist_mb": args.l2_persist_mb, "overlap_streams": bool(args.overlap), "isa_intv": args.isa_intv, "lit_ctas_per_sm": args.lit_ctas,
                       "library_tag": os.environ.get("COMPSEED_LIB_TAG", ""),
                       "parallelism": f"index replicated x{world}, reads sharded in contiguous blocks, host gather, no collective"},
            "parity": parity, "index_verify": index_verify,
            "occ_lookups_per_s_executed": occ_exec_per_read * value, "occ_lookups_per_read_executed": occ_exec_per_read,
            "occ_lookups_per_s_logical": ref_work["occ_lookups_per_s_logical"] if ref_work else None,
            "mems_per_read": n_mems / n_reads, "seeds_per_read": n_seeds / n_reads,
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": e2e, "host_link": host_link, "gpu_launches": int(dev_launches + e2e_launches),
            "gpu_launches_what": "kernel launches counted at the launch sites of the library (cs_ctx_launches): %d in the device-resident timed region "
                                 "(%d steps), %d in the host-buffer timed region" % (dev_launches, args.steps, e2e_launches),
            "clocks": clocks, "roofline": roofline, "reference_work_equivalent": ref_work, "cpu_baseline": cpu_baseline,
            "extension_stage": extension, "counters": counters, "setup_s": setup_s}
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()
    bad = (parity is not None and not parity["equal"]) or (index_verify is not None and not index_verify["ok"])
    if extension and "parity" in extension and not extension["parity"]["equal"]:
        bad = True
    if bad:
        sys.stderr.write("PARITY / INDEX CHECK FAILED: %s %s\n" % (parity, index_verify))
    return 3 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
