#!/usr/bin/env python
"""bench.py — Gbases/s of subject scanned by the blastn preliminary-search hot path.

Workload at N=1: BASELINE.json configs[1] — megablast, 1 000 x 1 kb synthetic queries (80 % planted,
2 % substitutions, half reverse-complemented) vs a 250 Mb synthetic DB (one chr1-sized sequence,
split at MAX_DBSEQ_LEN like the reference does).  At N>1 every rank owns one volume of the same
shape (volume sharding, weak scaling, no collective on the data path).

A "step" = one pass of the whole preliminary stage (scan -> mini-extension -> diagonal/ungapped ->
gapped score-only -> host replay -> E-values) of one query batch over one resident volume.
  value : subject bases scanned / second, volume + query tables already resident in HBM
  e2e   : same, through the host-buffer entry point (H2D of the packed volume and the query
          tables from pinned memory, search, D2H of results inside the timed region)
  roofline : scan kernel alone, algorithmic bytes = 0.25 B per subject base
  cpu_baseline : the reference engine (oracle/_ref, compiled from the reference's own sources)
          on this box's host cores, same inputs

`--impl reference` times the reference's own CPU implementation of the path instead.
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

from gblastn_b200 import synth  # noqa: E402

METRIC = "Gbases/s subject scanned (megablast) at 1/2/4/8 B200; HSPs bit-exact vs blastn"
UNIT = "Gbases/s"
N_QUERIES, QUERY_LEN, DB_BASES = 1000, 1000, 250_000_000
WORKLOAD = "megablast: 1000x1kb synthetic queries vs 250Mb synthetic DB (BASELINE configs[1])"
N_SETS = 4          # (volume, query batch) sets a rank cycles through: step i searches set i % N_SETS
L2_NOTE = ("inputs larger than L2: consecutive steps search different (volume, query batch) sets, 4 sets of 62.5 MB packed "
           "subject + ~20 MB of touched table data per GPU, cycled; no flush inside the timed region")


def bench_config(n_vol):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm."""
    return {"workload": WORKLOAD, "volumes": n_vol, "l2": L2_NOTE}


# ---- the other named configurations of BASELINE.json (configs[2..4]) at their stated sizes -------------------
# C3 runs whole on one GPU; C4 and C5 are volume-sharded over 4 / 8 GPUs: a rank builds and searches ITS shard
# (every rank sees the whole query batch).  At N=1 the line also carries one shard of C4 and of C5, labelled so.
def named_config(name, shard=0):
    if name == "C3":      # blastn ws 11: 100 x 10 kb vs 1 Gb (10 x 100 Mb), 8 % substitutions + 1 % indels
        vol = synth.random_volume([100_000_000] * 10, seed=3)
        qs = synth.planted_queries(vol, 100, 10_000, seed=33, planted_frac=0.8, sub_rate=0.08, indel_rate=0.01)
        return dict(task="blastn", vol=vol, qs=qs, masks=None, db_length=1_000_000_000, db_num_seqs=10,
                    workload="blastn word_size 11: 100x10kb queries vs 1Gb synthetic DB (10x100Mb), 1 GPU (BASELINE configs[2])",
                    sample_oids=1)
    if name == "C4":      # megablast: 100 k x 150 bp reads vs 3 Gb = 4 volumes x (7 x ~107 Mb)
        vol = synth.random_volume([107_142_857] * 7, seed=40 + shard)
        qs = synth.planted_queries(vol, 100_000, 150, seed=44, planted_frac=0.8, sub_rate=0.02)
        return dict(task="megablast", vol=vol, qs=qs, masks=None, db_length=3_000_000_000, db_num_seqs=28,
                    workload="megablast: 100kx150bp short reads vs 3Gb synthetic DB, volume-sharded 4 GPUs: "
                             "one shard = 7x107Mb (BASELINE configs[3])", sample_oids=7, ref_threads=7)
    if name == "C5":      # megablast + DUST: 1000 x 5 kb vs 20 Gb nt-like = 8 volumes x 2.5 Gb, log-normal lengths
        rng = np.random.default_rng(50 + shard)
        lens = np.clip(np.exp(rng.normal(np.log(2000.0), 1.2, size=625_000)).astype(np.int64), 30, 10_000_000)
        lens = lens[: int(np.searchsorted(np.cumsum(lens), 2_500_000_000)) + 1]
        vol = synth.random_volume(lens, seed=50 + shard)
        qs = synth.planted_queries(vol, 1000, 5000, seed=55, planted_frac=0.8, sub_rate=0.02, indel_rate=0.002)
        qs = synth.add_low_complexity(qs, seed=56, frac=0.3)
        from gblastn_b200 import engine as _E
        # blastn -dust yes (task default 20 64 1): the batch is masked on the device, the host routine runs beside it
        _E.dust_mask_batch(qs[:4])
        t0 = time.perf_counter()
        masks = _E.dust_mask_batch(qs)
        t_dev = time.perf_counter() - t0
        t0 = time.perf_counter()
        host_masks = [_E.dust_mask(q) for q in qs]
        t_host = time.perf_counter() - t0
        dust = {"device_ms": round(1e3 * t_dev, 3), "host_ms_1core": round(1e3 * t_host, 3), "identical": masks == host_masks,
                "intervals": int(sum(len(m) for m in masks)), "masked_bases": int(sum(b - a + 1 for m in masks for a, b in m))}
        return dict(task="megablast", vol=vol, qs=qs, masks=masks, dust=dust, db_length=int(8 * vol.total_bases), db_num_seqs=int(8 * vol.n_seqs),
                    workload="megablast + DUST: 1000x5kb queries (30 % with low-complexity inserts) vs 20Gb nt-like DB, "
                             "volume-sharded 8 GPUs: one shard = 2.5Gb of log-normal length sequences (BASELINE configs[4])",
                    sample_oids=10 ** 9, ref_threads=1)      # one thread: the hit lists overflow, low_score is order-dependent
    raise KeyError(name)


def run_named_config(name, shard, engine, setup, torch, steps=3, with_reference=True):
    """One named configuration (or one shard of it) on the current device: stage times, throughput, parity of a
    bounded sample against the reference engine, the reference's own speed on that sample."""
    t_gen = time.perf_counter()
    w = named_config(name, shard)
    vol, qs = w["vol"], w["qs"]
    t_gen = time.perf_counter() - t_gen
    kw = {}
    if w["task"] == "megablast":
        kw["device_lookup"] = 1
    t0 = time.perf_counter()
    s = setup.Setup(qs, task=w["task"], db_length=w["db_length"], db_num_seqs=w["db_num_seqs"], masks=w["masks"], **kw)
    V = engine.Volume(vol, device=0)
    Q = engine.Query(s.batch)
    t_load = time.perf_counter() - t0
    out = {"name": name, "workload": w["workload"], "shard": shard, "subject_bases": int(vol.total_bases),
           "subjects": int(vol.n_seqs), "query_batches": 1,
           "query_bases": int(sum(len(q) for q in qs)), "lut": f"lut {s.batch.lut_word_length} / stride {s.batch.scan_step}",
           "seconds_generate": round(t_gen, 2), "seconds_setup_and_load": round(t_load, 2)}
    if w.get("dust"):
        out["dust"] = w["dust"]
    try:
        engine.prelim_search(V, Q)                     # warm-up: buffers, chunk table
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms, stage = [], {"ms_scan": 0.0, "ms_extend": 0.0, "ms_gapped": 0.0, "ms_host": 0.0}
        g = None
        for _ in range(steps):
            torch.cuda.synchronize()
            ev0.record()
            g = engine.prelim_search(V, Q)
            ev1.record()
            ev1.synchronize()
            ms.append(ev0.elapsed_time(ev1))
            for k in stage:
                stage[k] += g["stats"][k] / steps
        st = g["stats"]
        out.update({"ms_per_pass": float(np.mean(ms)), "gbases_per_s": vol.total_bases / (np.mean(ms) * 1e-3) / 1e9,
                    "stage_ms": {k: round(v, 3) for k, v in stage.items()}, "hsps": int(g["hsps"].size),
                    "lookup_hits": int(st["lookup_hits"]), "init_hsps": int(st["good_init_extends"]),
                    "gap_extensions": int(st["gap_extensions"]), "kernel_launches_per_pass": int(st["kernel_launches"])})
        scan_ms, scan_bases, _ = engine.bench_scan(V, Q, 3)
        peak, _ = peak_hbm()
        out["scan_kernel"] = {"ms": scan_ms, "algorithmic_GBs": scan_bases * 0.25 / (scan_ms * 1e-3) / 1e9,
                              "frac_of_hbm_peak": scan_bases * 0.25 / (scan_ms * 1e-3) / 1e9 / peak}
        if with_reference:
            from oracle import refdriver as R, portdriver as P
            if R.available():
                # bounded sample: the first k subjects of the shard, searched alone by the reference with the
                # effective search space of the whole database (db_length / db_num_seqs), against the GPU lists
                # of the same OID range
                k = min(int(w["sample_oids"]), vol.n_seqs)
                sub = synth.Volume(vol.packed, vol.byte_off[:k], vol.seq_len[:k])
                cores = os.cpu_count() or 1
                threads = max(1, min(cores, k, int(w.get("ref_threads", 1))))
                cfg = R.default_config(w["task"], db_length=w["db_length"], db_num_seqs=w["db_num_seqs"], num_threads=threads)
                t0 = time.perf_counter()
                r = R.search(qs, sub, cfg, masks=w["masks"])
                wall = time.perf_counter() - t0
                gs = engine.prelim_search(V, Q, 0, k)
                # with several reference threads the per-thread lists come back in OID order like ours
                same = bool(r["status"] == 0 and np.array_equal(P.final_table(gs["hsps"]), r["final"]))
                out["parity_vs_reference"] = {"identical": same, "sample": f"{'all' if k == vol.n_seqs else 'first'} {k} subject(s) of the shard = "
                                              f"{int(sub.seq_len.astype(np.int64).sum())} bases, all queries",
                                              "hsps": int(r["final"].shape[0])}
                out["cpu_baseline"] = {"gbases_per_s": float(sub.seq_len.astype(np.int64).sum()) / r["seconds_prelim"] / 1e9,
                                       "cores": threads, "kind": "reference", "seconds": round(r["seconds_prelim"], 2),
                                       "wall_seconds": round(wall, 2)}
    finally:
        Q.free(); V.free(); s.free()
    return out


def make_workload(rank: int, k: int = 0):
    """Set k of rank `rank`: a 250 Mb volume and the 1000 x 1 kb queries planted in it (80 % planted, 2 % substitutions)."""
    vol = synth.random_volume([DB_BASES], seed=2 + 1000 * rank + 100 * k)
    qs = synth.planted_queries(vol, N_QUERIES, QUERY_LEN, seed=22 + 1000 * rank + 100 * k, planted_frac=0.8,
                               sub_rate=0.02, rc_frac=0.5)
    return vol, qs


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def scan_traffic():
    """dram bytes per scan launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz, self.source = [], set(), None, None
        self._stop = threading.Event()
        self.index = index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process (nvidia_ml_py): a query costs microseconds and spawns nothing; `nvidia-smi` every
        # 100 ms per rank was a measurable disturbance of sub-millisecond steps.  Falls back to the CLI.
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                N.nvmlDeviceGetCurrentClocksThrottleReasons
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            self.source = "nvml"
            while not self._stop.is_set():
                self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                self._stop.wait(0.05)
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip().splitlines()
                if out:
                    f = [x.strip() for x in out[0].split(",")]
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                        "sw_power_cap"), f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.25)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": self.source}


def bind_near_gpu(bus_id: str):
    """Run this rank on the cores of the NUMA node its GPU hangs off (host replay, pinned buffers and the
    launch path then stay on one socket).  Best effort: returns the node or None."""
    try:
        dev = "/sys/bus/pci/devices/" + bus_id.lower()
        if not os.path.isdir(dev) and len(bus_id.split(":")[0]) == 8:      # NVML prints an 8-digit domain
            dev = "/sys/bus/pci/devices/" + bus_id.lower()[4:]
        node = int(open(dev + "/numa_node").read())
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def _reference_worker(job):
    """One rank's share of the reference arm: `steps` timed preliminary searches, step i on set i % N_SETS."""
    r, warmup, steps = job
    from oracle import refdriver as R
    sets = [make_workload(r, k) for k in range(min(N_SETS, max(1, steps)))]
    cfg = R.default_config("megablast", num_threads=1)
    times = []
    for i in range(warmup + steps):
        vol, qs = sets[i % len(sets)]
        res = R.search(qs, vol, cfg)
        assert res["status"] == 0
        if i >= warmup:
            times.append(res["seconds_prelim"])
    return times


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref) on the host cores.

    Same configuration as the GPU arm at this N: N volumes, each searched with its own query batch.
    The reference parallelises over subject sequences (CPrelimSearchThread pulls OID ranges), so a
    volume that is ONE 250 Mb sequence keeps one thread busy; the N volumes run concurrently, one
    process each, which is every thread the reference can use on this workload."""
    if rank != 0:
        return
    from oracle import refdriver as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libblastref.so was not built"}))
        return
    n_vol = max(1, args.gpus)
    cores = os.cpu_count() or 1
    jobs = [(r, args.warmup, args.steps) for r in range(n_vol)]
    if n_vol == 1:
        per_vol = [_reference_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(n_vol, cores)) as pool:
            per_vol = pool.map(_reference_worker, jobs)
    total = max(float(sum(t)) for t in per_vol)          # the job ends when its slowest volume does
    steps = len(per_vol[0])
    value = n_vol * DB_BASES * steps / total / 1e9
    threads = min(n_vol, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": bench_config(n_vol),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"full workload per step: {n_vol} volume(s) searched concurrently, one thread each "
                                   f"(the reference parallelises over subject sequences; a volume is one sequence), "
                                   f"{cores} host cores present"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gblastn_b200", choices=["gblastn_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 blocks (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gblastn_b200 import engine, setup, abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = None
    if world > 1:
        try:
            import pynvml as N
            N.nvmlInit()
            # NVML enumerates like CUDA when CUDA_DEVICE_ORDER=PCI_BUS_ID; match by UUID to be safe
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            for i in range(N.nvmlDeviceGetCount()):
                h = N.nvmlDeviceGetHandleByIndex(i)
                u = N.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u or u.replace("GPU-", "") == uuid:
                    bus = N.nvmlDeviceGetPciInfo(h).busId
                    numa = bind_near_gpu(bus.decode() if isinstance(bus, bytes) else bus)
                    break
        except Exception:
            numa = None
        if numa is None:
            # no NUMA information for the GPU (single-node hosts report -1): give every rank its own slice of the
            # cores instead, so that the ranks' launch / replay / traceback threads never share a core
            try:
                cpus = sorted(os.sched_getaffinity(0))
                per = len(cpus) // world
                if per >= 2 and not os.environ.get("BN_BENCH_NO_BIND"):
                    os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]))
            except Exception:
                pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N_SETS (volume, query batch) sets per rank; the packed volumes live in pinned host memory (source of the e2e H2D copies)
    sets = [make_workload(rank, k) for k in range(N_SETS)]
    for vol_k, _ in sets:
        pinned = torch.empty(vol_k.packed.shape[0], dtype=torch.uint8).pin_memory()
        pinned.numpy()[:] = vol_k.packed
        vol_k.packed = pinned.numpy()
    vol, qs = sets[0]

    engine.init(0, [local_rank])
    # device_lookup: the megablast table is filled on the GPU at bn_query_load (s_FillContigMBTable
    # semantics), so a step's H2D is the packed volume + the query bytes, not the 4^lut-entry hashtable
    setups = [setup.Setup(q_k, task="megablast", db_length=v_k.total_bases, db_num_seqs=v_k.n_seqs, device_lookup=1)
              for v_k, q_k in sets]
    s = setups[0]
    Vs = [engine.Volume(v_k, device=0) for v_k, _ in sets]
    Qs = [engine.Query(st.batch) for st in setups]
    V, Q = Vs[0], Qs[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    def resident_jobs(n):
        return [{"volume": Vs[i % N_SETS], "query": Qs[i % N_SETS]} for i in range(n)]

    def host_jobs(n):
        return [{"host_volume": sets[i % N_SETS][0], "batch": setups[i % N_SETS].batch} for i in range(n)]

    # ---- resident-input throughput: K steps = K jobs through the job pipeline (bn_prelim_search_jobs) ----------
    # Every job is a complete preliminary search (scan ... gapped on the device, replay / E-values on the host); the
    # pipeline queues job k+1's kernels before it waits for job k and replays finished jobs on a worker thread.
    results = engine.prelim_search_jobs(resident_jobs(max(args.warmup, N_SETS, 12)))      # every lane of the pipeline has run twice
    per_set = results[:N_SETS]                       # one result per set, for the parity check below
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"ms_scan": 0.0, "ms_extend": 0.0, "ms_gapped": 0.0, "ms_host": 0.0}
    # the timed region is the C-ABI call alone: when it returns, every job's results are in host memory (the ctypes
    # job array is built before, the numpy views of the results are made after)
    timed = engine.JobBatch(resident_jobs(args.steps))
    barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        ev0.record()
        timed.run()
        ev1.record()
        ev1.synchronize()
        total_ms_wall = 1e3 * (time.perf_counter() - t0)
        total_ms = float(ev0.elapsed_time(ev1))
        results = timed.results()
        launches = sum(r["stats"]["kernel_launches"] for r in results)
        for r in results:
            for k in stage:
                stage[k] += r["stats"][k]
        g = results[0]
        # ---- end to end: the same K steps with every input in HOST memory — per job the packed volume (pinned) and
        # the query batch cross PCIe, the lookup table is filled on the device, the results come back ----------------
        engine.prelim_search_jobs(host_jobs(max(args.warmup, 12)))      # every lane of the pipeline has sized its arena
        timed = engine.JobBatch(host_jobs(args.steps))
        barrier()
        ev0.record()
        timed.run()
        ev1.record()
        ev1.synchronize()
        total_e2e_ms = float(ev0.elapsed_time(ev1))
        e2e_results = timed.results()
        ge = e2e_results[0]
        e2e_same = all(a["hsps"].tobytes() == b["hsps"].tobytes() for a, b in zip(results, e2e_results))
        # ---- one blocking call per step (bn_prelim_search / bn_prelim_search_host), L2 flushed in between: what a
        # caller without a job stream sees -------------------------------------------------------------------------
        single_ms, single_e2e_ms = [], []
        for i in range(min(args.steps, 10)):
            l2_flush()
            ev0.record()
            engine.prelim_search(V, Q)
            ev1.record()
            ev1.synchronize()
            single_ms.append(ev0.elapsed_time(ev1))
        for i in range(min(args.steps, 5)):
            l2_flush()
            ev0.record()
            engine.prelim_search_host(s.batch, vol, device=0)
            ev1.record()
            ev1.synchronize()
            single_e2e_ms.append(ev0.elapsed_time(ev1))
        # ---- scan kernel alone (roofline) -----------------------------------------------------------
        engine.bench_scan(V, Q, 3)
        scan_ms, scan_bases, survivors = engine.bench_scan(V, Q, max(args.steps, 10))
    n_e2e = args.steps
    t = torch.tensor([total_ms, total_e2e_ms, scan_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mine = torch.tensor([total_ms / args.steps, total_e2e_ms / n_e2e, stage["ms_scan"] / args.steps,
                         stage["ms_extend"] / args.steps, stage["ms_gapped"] / args.steps, stage["ms_host"] / args.steps,
                         -1.0 if numa is None else float(numa)], dtype=torch.float64, device="cuda")
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    total_ms, total_e2e_ms, scan_ms_max = [float(x) for x in t.tolist()]
    bases_per_step = int(g["stats"]["subject_bases_scanned"])
    value = world * bases_per_step * args.steps / (total_ms * 1e-3) / 1e9
    e2e_value = world * bases_per_step * n_e2e / (total_e2e_ms * 1e-3) / 1e9

    b = s.batch
    h2d = int(vol.packed.shape[0] + (b.concat_len + 2) + 32 * b.num_contexts + 8 * b.n_lookup_segments + 2048)
    d2h = int(ge["hsps"].size * abi.HSP_DTYPE.itemsize + g["stats"]["good_init_extends"] * (32 + 32) + 64)

    peak, peak_kind = peak_hbm()
    achieved = scan_bases * 0.25 / (scan_ms * 1e-3) / 1e9      # GB/s of algorithmic bytes
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(world),
        "details": {"lut": f"MB lut {b.lut_word_length} / stride {b.scan_step}, filled on the device",
                    "hsps_per_step": int(g["hsps"].size)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": scan_traffic(), "kernel": "bn::scan_kernel_staged<0,17,1>",
                     "peak_kind": peak_kind, "ms_per_launch": scan_ms,
                     "algorithmic_bytes_per_launch": scan_bases * 0.25},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
        "ms_per_step_wall": total_ms_wall / args.steps,
        "timing": "CUDA events around the K steps (K complete searches through the job pipeline, the last host replay "
                  "included); max over ranks.  stage_ms_per_step are per-search device spans; consecutive searches overlap",
        "single_call": {"ms_per_step": float(np.mean(single_ms)), "gbases_per_s": bases_per_step / (np.mean(single_ms) * 1e-3) / 1e9,
                        "e2e_ms_per_step": float(np.mean(single_e2e_ms)),
                        "e2e_gbases_per_s": bases_per_step / (np.mean(single_e2e_ms) * 1e-3) / 1e9,
                        "note": "one blocking bn_prelim_search / bn_prelim_search_host call per step, L2 flushed between steps (this rank)"},
        "e2e_equals_resident": bool(e2e_same),
        "clocks": clk.summary(),
        "ranks": [dict(zip(("ms_per_step", "e2e_ms_per_step", "ms_scan", "ms_extend", "ms_gapped", "ms_host", "numa"),
                           [round(float(x), 4) for x in r.tolist()])) for r in per_rank],
    }

    # ---- small query batch on the same volume (BASELINE configs[0]'s query shape: ONE 10 kb query; N=1 only): the scan
    # goes through scan_kernel_filtered (presence filter of the table in shared memory), timed beside the queue-driven
    # kernel on the same batch; its fraction of the HBM peak is reported here, separately from `roofline` ------------------
    if rank == 0 and world == 1:
        try:
            qs1 = synth.planted_queries(vol, 1, 10_000, seed=5, planted_frac=1.0, sub_rate=0.02, rc_frac=0.5)
            s1 = setup.Setup(qs1, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
            sb = {"workload": "megablast: 1x10kb synthetic query (BASELINE configs[0] query shape) vs the 250Mb volume of configs[1]",
                  "lut": f"MB lut {s1.batch.lut_word_length} / stride {s1.batch.scan_step}"}
            outs = {}
            for mode in ("filtered", "queue"):
                if mode == "queue":
                    os.environ["BN_FILT_MAX"] = "0"
                else:
                    os.environ.pop("BN_FILT_MAX", None)
                Q1 = engine.Query(s1.batch)
                engine.bench_scan(V, Q1, 3)
                ms1, bases1, _ = engine.bench_scan(V, Q1, max(args.steps, 10))
                engine.prelim_search(V, Q1)
                l2_flush()
                ev0.record()
                outs[mode] = engine.prelim_search(V, Q1)
                ev1.record()
                ev1.synchronize()
                sb[mode] = {"scan_ms": ms1, "algorithmic_GBs": bases1 * 0.25 / (ms1 * 1e-3) / 1e9,
                            "frac_of_hbm_peak": bases1 * 0.25 / (ms1 * 1e-3) / 1e9 / peak,
                            "search_ms": float(ev0.elapsed_time(ev1)),
                            "kernel": "bn::scan_kernel_filtered" if mode == "filtered" else "bn::scan_kernel_staged"}
                Q1.free()
            os.environ.pop("BN_FILT_MAX", None)
            sb["identical"] = bool(outs["filtered"]["hsps"].tobytes() == outs["queue"]["hsps"].tobytes() and
                                   outs["filtered"]["stats"]["lookup_hits"] == outs["queue"]["stats"]["lookup_hits"])
            sb["hsps"] = int(outs["filtered"]["hsps"].size)
            sb["lookup_hits"] = int(outs["filtered"]["stats"]["lookup_hits"])
            if not args.no_cpu_baseline:
                from oracle import refdriver as R, portdriver as P
                if R.available():
                    r1 = R.search(qs1, vol, R.default_config("megablast", num_threads=1))
                    sb["parity_vs_reference"] = bool(np.array_equal(P.final_table(outs["filtered"]["hsps"]), r1["final"]))
                    sb["reference_ms_1core"] = 1e3 * r1["seconds_prelim"]
            s1.free()
            line["small_batch"] = sb
        except Exception as e:
            line["small_batch"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline (reference engine on the host cores), rank 0 at N=1 only --------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refdriver as R, portdriver as P
            if R.available():
                cores = os.cpu_count() or 1
                threads = max(1, min(cores, vol.n_seqs))
                cfg = R.default_config("megablast", num_threads=threads)
                secs, reps = 0.0, 0
                t_begin = time.perf_counter()
                while reps < 3 or (time.perf_counter() - t_begin < 10.0 and reps < 40):
                    r = R.search(qs, vol, cfg)
                    secs += r["seconds_prelim"]
                    reps += 1
                line["cpu_baseline"] = {
                    "value": DB_BASES * reps / secs / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
                    "sample": f"{reps} repeats of the full workload (prelim stage only); the reference "
                              f"parallelises over subject sequences, DB has {vol.n_seqs} -> {threads} thread"}
                same_all = bool(np.array_equal(P.final_table(per_set[0]["hsps"]), r["final"]))
                for k in range(1, N_SETS):
                    rk = R.search(sets[k][1], sets[k][0], cfg)
                    same_all = same_all and bool(np.array_equal(P.final_table(per_set[k]["hsps"]), rk["final"]))
                line["parity_vs_reference"] = same_all
                # ---- the stage after the path, outside every timed region above: the same
                # step's preliminary lists through bn_traceback_search, next to the reference's traceback stage
                try:
                    xf = s.gap_x_dropoff_final()
                    engine.traceback_search(V, Q, xf, g["hsps"])
                    tb_ms = []
                    for _ in range(5):
                        t0 = time.perf_counter()
                        tb, tb_ops = engine.traceback_search(V, Q, xf, g["hsps"])
                        tb_ms.append(1e3 * (time.perf_counter() - t0))
                    rt = R.search(qs, vol, R.default_config("megablast", taps=R.TAP_TRACEBACK, prelim_only=0))
                    w = rt["tb_final"]
                    same = tb.shape[0] == w.shape[0] and all(
                        np.array_equal(tb[c], w[:, k]) for k, c in enumerate(
                            ("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")))
                    line["traceback_stage"] = {
                        "ms": float(np.median(tb_ms)), "hsps": int(tb.shape[0]), "edit_ops": int(tb_ops.shape[0]),
                        "reference_ms_1core": 1e3 * rt["seconds_traceback"], "identical_to_reference": bool(same),
                        "note": "not part of value / e2e; wall time of bn_traceback_search on one step's lists"}
                except Exception as e:
                    line["traceback_stage"] = {"ms": None, "note": f"failed: {e}"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref/libblastref.so not present on this box"}
        except Exception as e:  # the baseline must never take the bench down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": f"failed: {e}"}
    for x in Qs + Vs + setups:
        x.free()
    del flush
    torch.cuda.empty_cache()

    # ---- the other named configurations at their stated sizes (not part of value / e2e) ---------------------------
    # N = 1: C3 whole, plus ONE shard each of C4 and C5 measured on this GPU; N = 4: C4, N = 8: C5, one shard per rank.
    if not args.no_configs:
        blocks = []
        plan = {1: [("C3", 0), ("C4", 0), ("C5", 0)], 4: [("C4", rank)], 8: [("C5", rank)]}.get(world, [])
        for name, shard in plan:
            try:
                blk = run_named_config(name, shard, engine, setup, torch, steps=3,
                                       with_reference=(rank == 0 and not args.no_cpu_baseline))
            except Exception as e:      # a named block must never take the headline down
                blk = {"name": name, "shard": shard, "error": f"{type(e).__name__}: {e}"}
            if world > 1 and "ms_per_pass" in blk:
                # the sharded configuration as a whole: all shards run concurrently, the job ends with the slowest
                tt = torch.tensor([blk["ms_per_pass"], float(blk["subject_bases"]), float(blk["hsps"])],
                                  dtype=torch.float64, device="cuda")
                mx = tt.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                sm = tt.clone()
                dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                blk["all_shards"] = {"gpus": world, "ms_per_pass_max_over_ranks": float(mx[0]),
                                     "subject_bases": int(sm[1]), "hsps": int(sm[2]),
                                     "gbases_per_s": float(sm[1]) / (float(mx[0]) * 1e-3) / 1e9}
            elif world > 1:
                dist.barrier()
            blocks.append(blk)
        line["configs"] = blocks
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
