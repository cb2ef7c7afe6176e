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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
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
    ap.add_argument("--e2e-batch", type=int, default=1 << 20, help="reads per pipelined batch on the host-buffer path")
    ap.add_argument("--e2e-slots", type=int, default=3, help="pipelined batches in flight on the host-buffer path")
    ap.add_argument("--cpu-sample", type=int, default=300_000, help="reads of the same workload timed on the host cores")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--probe", action="store_true", help="also measure the random 32-byte sector gather peak")
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
    or, when oracle/_ref is absent, our C port, on the first n_sample reads of the workload."""
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
        res = cs
    else:
        oi = O.OracleIndex.from_arrays(host_idx["primary"], host_idx["L2"], host_idx["seq_len"], host_idx["bwt"], host_idx["sa"], host_idx["sa_intv"])
        res = oi.seed(b, o, n_threads=threads)
        out.update(kind="port", value=n / res.seconds, unit="reads/s", seconds=res.seconds)
    return out, res, n


def work_counters(host_idx, bases, off, n_sample: int, threads: int):
    """Implementation-independent work per read from the instrumented C port."""
    from oracle import oracle_py as O
    n = min(n_sample, off.shape[0] - 1)
    oi = O.OracleIndex.from_arrays(host_idx["primary"], host_idx["L2"], host_idx["seq_len"], host_idx["bwt"], host_idx["sa"], host_idx["sa_intv"])
    r = oi.seed(bases[:int(off[n])], off[:n + 1], n_threads=threads)
    c = r.counters
    return {k: c[k] / n for k in c}, r, n


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
    idx = cs.FMIndex.build(ref, device=local_rank, sa_intv=args.sa_intv)
    del ref
    opt = cs.SeedOpt()
    setup_s = time.time() - t_setup

    # ---------------------------------------------------------------------------------------
    if args.impl == "reference":
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
    # our arm.  (1) device-resident: all reads of the step staged in HBM once
    max_mems, max_seeds = n_reads * 14, n_reads * 20
    ctx = cs.SeedContext(idx, n_reads, int(off[-1]), args.read_len, max_mems, max_seeds, 1)
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
    sampler.mark_begin()
    t0 = time.perf_counter()
    dev_ms, seed_ms, coll_ms, sa_ms, r3_ms, fast_ms, walk_ms = 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0
    for _ in range(args.steps):
        ctx.run_staged(0, opt)
        last = ctx.wait_device(0)
        seed_ms += last.kernel_ms[4]; r3_ms += last.kernel_ms[5]; fast_ms += last.kernel_ms[6]; walk_ms += last.kernel_ms[7]
        coll_ms += last.kernel_ms[1]
        sa_ms += last.kernel_ms[2]
        dev_ms += last.kernel_ms[0] + last.kernel_ms[1] + last.kernel_ms[2]   # CUDA events on the launching stream
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.mark_end()
    clocks = sampler.stop()
    counters = last.counters
    n_mems, n_seeds = last.n_mems_device, last.n_seeds_device
    # parity spot check at full size: size-independent properties
    got = ctx.fetch(0)
    starts = (got.mems[:, 3] >> np.uint64(32)).astype(np.int64)
    ends = (got.mems[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
    assert got.mem_off[-1] == n_mems and got.seed_off[-1] == n_seeds
    assert np.all(ends - starts >= opt.min_seed_len) and np.all(ends <= args.read_len) and np.all(got.mems[:, 2] >= 1)
    assert np.all(got.rbeg >= 0) and np.all(got.rbeg < idx.seq_len)
    info = got.mems[:, 3]
    same_read = np.ones(info.shape[0], dtype=bool)
    same_read[got.mem_off[1:-1][got.mem_off[1:-1] < info.shape[0]]] = False
    assert np.all((info[1:] >= info[:-1]) | ~same_read[1:]), "mems not sorted by info inside a read"
    ctx.close()

    if use_dist:
        t = torch.tensor([dev_ms, wall_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    else:
        dev_ms_max, wall_ms_max = dev_ms, wall_ms
    total_reads = n_reads * world * args.steps
    value = total_reads / (dev_ms_max * 1e-3)

    # (2) end to end through the C-ABI with HOST buffers: pipelined batches, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        bs = min(args.e2e_batch, n_reads)
        n_slots = args.e2e_slots
        cs.host_register(bases)   # the reads sit in page-locked host memory, as the bench contract asks
        ectx = cs.SeedContext(idx, bs, bs * args.read_len, args.read_len, bs * 14, bs * 20, n_slots)
        starts_b = list(range(0, n_reads, bs))

        def one_pass():
            h2d = d2h = 0
            inflight = []
            nxt = 0

            def submit(bi):
                s = starts_b[bi]
                e = min(n_reads, s + bs)
                o = off[s:e + 1] - off[s]
                ectx.submit(bi % n_slots, bases[int(off[s]):int(off[e])], o, opt)
                return int(off[e]) - int(off[s]) + 4 * (e - s + 1)

            while nxt < len(starts_b) and len(inflight) < n_slots:
                h2d += submit(nxt)
                inflight.append(nxt)
                nxt += 1
            tot_m = tot_s = 0
            while inflight:
                bi = inflight.pop(0)
                r = ectx.wait(bi % n_slots, copy=False)
                tot_m += int(r.mem_off[-1])
                tot_s += int(r.seed_off[-1])
                d2h += 4 * r.mem_off.shape[0] * 2 + 32 * int(r.mem_off[-1]) + 8 * int(r.seed_off[-1])
                if nxt < len(starts_b):
                    h2d += submit(nxt)
                    inflight.append(nxt)
                    nxt += 1
            return h2d, d2h, tot_m, tot_s

        for _ in range(max(1, args.warmup - 1)):
            one_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            h2d, d2h, tot_m, tot_s = one_pass()
        barrier()
        e_ms = (time.perf_counter() - t0) * 1e3
        assert tot_m == n_mems and tot_s == n_seeds, "host-buffer path disagrees with the device-resident path"
        if use_dist:
            t = torch.tensor([e_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t[0])
        e2e = {"value": total_reads / (e_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "batch_reads": bs, "slots": n_slots, "timing": "host wall clock between device syncs; reads in page-locked host memory, results read back into the slots' pinned buffers"}
        ectx.close()
        cs.host_unregister(bases)

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return 0

    # (3) roofline of the dominant kernel (k_seed) + same-run CPU baseline (rank 0, N=1 only for the CPU leg)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    cpu_baseline, per_read = None, None
    if not args.no_cpu and world == 1:
        host_idx = idx.download(sa_intv=32)
        cpu_baseline, _, _ = cpu_seed_sample(host_idx, bases, off, args.cpu_sample, threads)
        per_read, _, n_cnt = work_counters(host_idx, bases, off, min(args.cpu_sample, 100_000), threads)
        del host_idx
    # E = the reference's bwt_extend call count per read (logical work).  The device counters
    # are lower: the occurrence filter and the top-of-search table remove work without changing results.
    if per_read is not None:
        E = per_read["ext"]
        e2_ratio = per_read["ext2"] / per_read["ext"]
        S, A, M = per_read["lf"], per_read["sa"], per_read["mem"]
    else:
        E, e2_ratio, S, A, M = 847.6, 0.6, 31.0 * n_seeds / n_reads, n_seeds / n_reads, n_mems / n_reads
    # bytes of the seeding kernel = 64*(E+E2) + input bases + 32*M;  SA walk = 64*S + 16*A
    seed_bytes_per_read = 64.0 * E * (1.0 + e2_ratio) + args.read_len + 32.0 * M
    path_bytes_per_read = seed_bytes_per_read + 64.0 * S + 16.0 * A
    seed_ms_per_launch = (seed_ms + r3_ms) / args.steps      # k_seed + k_seed_r3 == mem_collect_intv
    achieved = seed_bytes_per_read * n_reads / (seed_ms_per_launch * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k_seed_traffic.json")))["dram_bytes_per_read"] * n_reads
    except Exception:
        pass
    # The same kernel seen from the memory system instead of from the reference's logical work: DRAM bytes ncu counted
    # for k_seed_fast (per read, from the committed capture) over its live duration, and its L2 read requests per
    # second next to the random-gather peak of this part (cs_probe_random_gather: 36-46 G loads/s).
    measured = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "k_seed_traffic.json")))
        fast_s = fast_ms / args.steps * 1e-3
        measured = {"kernel": "k_seed_fast (with k_pack_reads)", "dram_gb_per_s": tj["dram_bytes_per_read"] * n_reads / fast_s / 1e9,
                    "dram_frac_of_peak": tj["dram_bytes_per_read"] * n_reads / fast_s / 1e9 / peak,
                    "l2_read_requests_per_s": tj.get("l2_read_requests_per_read", 98.7) * n_reads / fast_s,
                    "random_gather_peak_loads_per_s": 37.9e9,
                    "what": "achieved/frac above count the reference's logical bytes (SURVEY 8d), which the result-neutral structures mostly "
                            "remove, hence frac > 1; these are the bytes and requests the dominant kernel really moves"}
    except Exception:
        pass
    roofline = {"kernel": "k_seed_fast + k_seed_walk + k_seed + k_seed_r3 (the three passes of mem_collect_intv)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "measured_traffic_view": measured,
                "traffic": traffic, "traffic_of": "k_seed_fast alone (the dominant kernel; ncu dram__bytes_read + write per read x reads of a launch, profiles/k_seed_traffic.json)",
                "peak_source": peak_src, "ms_per_launch": seed_ms_per_launch,
                "algorithmic_bytes_per_read": seed_bytes_per_read, "whole_path_bytes_per_read": path_bytes_per_read,
                "extends_per_read": E, "two_bucket_ratio": e2_ratio,
                "device_extends_per_read": counters["ext_queries"] / n_reads, "device_fm_extends_per_read": counters["ext_calls"] / n_reads,
                "fast_ms_per_launch": fast_ms / args.steps, "deferred_calls_per_read": counters.get("deferred_calls", 0) / n_reads,
                "kernel_share_of_step": {"k_seed_fast": fast_ms / dev_ms, "k_seed_walk": walk_ms / dev_ms, "k_seed": (seed_ms - fast_ms - walk_ms) / dev_ms, "k_seed_r3": r3_ms / dev_ms, "collect": coll_ms / dev_ms, "k_sa_resolve": sa_ms / dev_ms}}
    occ_per_read = 2.0 * E + S
    if args.probe:
        gb, gl = cs.probe_random_gather(local_rank, 4 << 30, 32, 1 << 28, 2)
        roofline["random_sector_peak"] = {"gloads_per_s": gl, "gb_per_s": gb, "what": "independent random 32-B loads over 4 GiB"}
        roofline["sector_reads_per_s"] = E * (1.0 + e2_ratio) * n_reads / (seed_ms_per_launch * 1e-3) / 1e9

    line = {"metric": "smem_seeding_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "sa_intv_device": args.sa_intv, "index_bytes_per_gpu": idx.device_bytes,
                       "result_neutral_structures": "dense SA, top-of-search k-mer table (depth <= 13), 2-bit occurrence filter (K <= 19), "
                                                    "2-bit text + sampled inverse SA for unique matches (DESIGN.md section 5)",
                       "l2_policy": "inputs larger than L2 (index %.1f GB, reads %.1f GB per step)" % (idx.device_bytes / 1e9, bases.nbytes / 1e9),
                       "parallelism": f"index replicated x{world}, reads sharded in contiguous blocks, host gather, no collective"},
            "occ_lookups_per_s": occ_per_read * value, "occ_lookups_per_read": occ_per_read,
            "mems_per_read": n_mems / n_reads, "seeds_per_read": n_seeds / n_reads,
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": e2e, "gpu_launches": 9 * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "counters": counters, "setup_s": setup_s}
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
