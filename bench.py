#!/usr/bin/env python3
"""bench.py -- MINRMS per-orientation search throughput on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4,5}]
                  [--scaling {weak,strong}] [--orientations M] [--impl reference]

Default workload (BASELINE.json configs[1], `--config 2`): the reference's own test pair --
hemoglobin chain A (141 CA) vs myoglobin (151 CA), tests/test_typical_usage.sh -- searched
UNCONSTRAINED: `minrms -fm 4 -HS -ir` = 20 424 orientations x 1 044 481 DP cells, all 140
alignment sizes refined by Kabsch for every orientation.  One "step" = one complete FirstPass
over the orientations of this rank, followed (N > 1) by the cross-GPU reduction of the per-size
records: an NCCL all_reduce(MIN) over the 64-bit keys in device memory, the winners' payload
from the owning ranks, one all_gather (minrms_b200/dist.py).

  --scaling weak    every rank searches its own copy of the list under its own ids (per-GPU
                    work fixed; the driver's SCALE run)
  --scaling strong  ONE list, partitioned over the ranks (dist.partition): what `minrms A B`
                    on N GPUs does; the merged table is checked against rank 0's own search
  --config 1        the test pair with the reference's test flags (-of 8.0,0.33 -minN 0.33
                    -max-rmsd 3.5): NW filter + early exits
  --config 3 / 4    synthetic 600x600 (-fm 3) / 2000x1500 (-of 8.0,0.33) pairs, seeded
  --config 5        all-vs-all of --chains synthetic chains (150-400 residues), pairs sharded
  --orientations M  search only the first M orientations of the list per step

Prints ONE JSON line (rank 0).  `value` = orientations/s with the orientation list already
resident in HBM; `e2e` = the same through the host-buffer C-ABI call (mr_search: H2D of the
list from pinned memory + D2H of the best table inside the timed region).  On the default
line also: `full_run` (orientation generation + FirstPass + Refine through the public API)
and `nw_filter` (the NW orientation filter of config 1 with its own roofline), and `alternates`
(the FirstPass variants behind -alt-o / -alt-r on the same workload).
`--impl reference` times the UNMODIFIED reference (oracle/_ref/fast, built by
oracle/build_ref.sh) on the host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CPU_CELLS_PER_S_PER_CORE = 0.22e9   # Dyn3d::Solve_IncrN of the reference, measured; used only where stated


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------- workloads
def workload(cfg):
    """pair, settings and the reference command line of a BASELINE.json configuration"""
    from minrms_b200 import Settings, load_ca_pdb
    from minrms_b200.search import interpret_num_matches
    from minrms_b200.synth import random_chain, related_chain
    if cfg in (1, 2):
        A = load_ca_pdb(os.path.join(GOLDEN, "hemoglobin.pdb"))
        B = load_ca_pdb(os.path.join(GOLDEN, "myoglobin.pdb"))
        pdbs = (os.path.join(GOLDEN, "hemoglobin.pdb") + ",*.A", os.path.join(GOLDEN, "myoglobin.pdb"))
        data = "fixture (CA traces of the reference's test pair, tests/golden/*.pdb)"
        if cfg == 2:
            return dict(cfg=2, A=A, B=B, pdbs=pdbs, data=data, settings=Settings(fm_fragment_size=4),
                        ref_args=["-fm", "4", "-HS", "-ir", "-R"],
                        name="hemoglobin(4hhb:A,141)xmyoglobin(1a6m,151) unconstrained: minrms -fm 4 -HS -ir (config 2)")
        k = interpret_num_matches(0.33, 141, 151)
        st = Settings(fm_fragment_size=4, n_min=k, dyn3d_max_rmsd=3.5, orientation_filter_nw=True,
                      orientation_filter_nw_max_dist=8.0, orientation_filter_nw_min_num_matches=k)
        return dict(cfg=1, A=A, B=B, pdbs=pdbs, data=data, settings=st,
                    ref_args=["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.5", "-ir", "-R"],
                    name="hemoglobin(4hhb:A,141)xmyoglobin(1a6m,151): minrms -fm 4 -of 8.0,0.33 -HS -minN 0.33 -max-rmsd 3.5 -ir (config 1)")
    if cfg == 3:
        A = random_chain(np.random.default_rng(3000), 600)
        B = related_chain(np.random.default_rng(3001), A, 600, noise=1.5)
        return dict(cfg=3, A=A, B=B, pdbs=None, data="synthetic (seeds 3000/3001, minrms_b200/synth.py)",
                    settings=Settings(fm_fragment_size=3), ref_args=["-fm", "3", "-HS", "-ir", "-R"],
                    name="synthetic 600x600, minrms -fm 3 -HS -ir, all sizes (config 3)")
    if cfg == 4:
        A = random_chain(np.random.default_rng(4000), 2000)
        B = related_chain(np.random.default_rng(4001), A, 1500, noise=1.0)
        k = interpret_num_matches(0.33, 2000, 1500)
        st = Settings(fm_fragment_size=4, orientation_filter_nw=True, orientation_filter_nw_max_dist=8.0,
                      orientation_filter_nw_min_num_matches=k)
        return dict(cfg=4, A=A, B=B, pdbs=None, data="synthetic (seeds 4000/4001, minrms_b200/synth.py)", settings=st,
                    ref_args=["-fm", "4", "-of", "8.0,0.33", "-HS", "-ir", "-R"],
                    name="synthetic 2000x1500, minrms -fm 4 -of 8.0,0.33 -HS -ir (config 4)")
    raise SystemExit("unknown --config %r" % cfg)


def ref_bench(w, procs, per_proc, repeat):
    """the reference's own Search(og,start,stop) on `procs` host processes (oracle/ref_bench.cc)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "fast", "ref_bench")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        pdbs = w["pdbs"]
        if pdbs is None:
            from minrms_b200.synth import write_ca_pdb
            write_ca_pdb(os.path.join(tmp, "a.pdb"), w["A"])
            write_ca_pdb(os.path.join(tmp, "b.pdb"), w["B"])
            pdbs = (os.path.join(tmp, "a.pdb"), os.path.join(tmp, "b.pdb"))
        out = subprocess.run([exe, str(procs), str(per_proc), str(repeat)] + w["ref_args"] + list(pdbs),
                             cwd=tmp, capture_output=True, text=True)
    res = []
    for line in out.stdout.splitlines():
        if line.startswith("ref_bench"):
            kv = dict(x.split("=") for x in line.split()[1:])
            if int(kv["failed"]) == 0:
                res.append((int(kv["orientations"]), float(kv["seconds"]), int(kv["total"])))
    return res


def ref_sample_size(w):
    """orientations per process and step for about a second of reference work per step"""
    m, nB = w["A"].shape[0], w["B"].shape[0]
    if w["cfg"] == 1:
        return 1100          # NW filter on all, DP on 15 %: the whole list in ~1 s on 16 cores
    cells = dp_cells_closed_form(m, nB, 1, min(m, nB))
    return max(1, int(CPU_CELLS_PER_S_PER_CORE / cells))


def dp_cells_closed_form(m, nB, n_min, n_max):
    """DP cells of one orientation: layer n covers rows n..min(m, n + wA - 1) and
    columns n..min(nB, n + wB - 1) with wA = m - n_min + 1, wB = nB - n_min + 1.  Plain Python, so that
    the reference arm does not load the CUDA library for it."""
    wA, wB = m - n_min + 1, nB - n_min + 1
    return sum((min(m, n + wA - 1) - n + 1) * (min(nB, n + wB - 1) - n + 1) for n in range(1, n_max + 1))


def common_config(w, args, count):
    """the `config` object: both arms print exactly this (the driver compares them); everything
    in it follows from the workload and the launch, nothing from what a run measured"""
    m, nB = w["A"].shape[0], w["B"].shape[0]
    st = w["settings"]
    n_min = int(getattr(st, "n_min", 1) or 1)
    n_max = min(m, nB)
    cells = dp_cells_closed_form(m, nB, n_min, n_max)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = args.scaling == "strong" and world > 1
    per_gpu = (count + world - 1) // world if strong else count
    return {"workload": w["name"], "config": w["cfg"], "scaling": args.scaling if world > 1 else "weak",
            "orientations_in_list": int(count), "orientations_per_gpu_per_step": int(per_gpu),
            "dp_cells_per_orientation": cells, "alignment_sizes": int(n_max - max(2, n_min) + 1),
            "l2": "no explicit flush: the traceback tables of a step (0.25 B per DP cell of every orientation that reaches "
                  "the DP, up to %.1f GB per GPU for this list) are written and re-read, >> 126 MB L2" % (0.25 * cells * per_gpu / 1e9),
            "parallelism": ("one list partitioned over %d ranks (contiguous ranges, global ids)" % world if strong else
                            "%d rank(s), each its own copy of the list under its own ids" % world) +
                           "; per step: NCCL all_reduce(MIN) of the 64-bit per-size keys in device memory (reduction_ms); "
                           "the winners' payload from the owning ranks + one all_gather once per search "
                           "(payload_ms_once_per_search; inside every step of e2e)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.config)
    m, nB = w["A"].shape[0], w["B"].shape[0]
    cores = os.cpu_count() or 1
    per_proc = ref_sample_size(w)
    if w["cfg"] in (3, 4):
        # one orientation costs minutes of a host core here (72 M / 1 689 M cells at ~0.2 GCUPS; at
        # config-4 size any survivor of the filter): no bounded sample exists
        print(json.dumps({"impl": "reference", "unavailable": "config %d: one orientation costs minutes of a host core "
                          "(DESIGN.md 3: CPU figures for configs 3-4 are extrapolated from the measured cell rate)" % w["cfg"]}))
        return
    res = ref_bench(w, cores, per_proc, args.warmup + args.steps)
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast/ref_bench not built"}))
        return
    total = res[0][2]
    res = res[args.warmup:]
    n = sum(r[0] for r in res)
    t = sum(r[1] for r in res)
    v = n / t
    cells = dp_cells_closed_form(m, nB, w["settings"].n_min, min(m, nB))
    count = min(total, args.orientations) if args.orientations > 0 else total
    sample = ("%d orientations/step (%d per process x %d processes) of the %d, Search(og,start,stop) of the unmodified "
              "reference, shipped flags -O3 -ffast-math" % (res[0][0], res[0][0] // cores, cores, total))
    line = {"metric": "orientations_per_sec", "value": v, "unit": "orientations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(res),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": w["data"], "impl": "reference", "dp_gcups": v * cells / 1e9,
            "config": common_config(w, args, count),
            "cpu_baseline": {"value": v, "unit": "orientations/s", "cores": cores, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": v, "unit": "orientations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------- config 5
def run_all_vs_all(args):
    """BASELINE config 5 shape: every unordered pair of --chains synthetic chains (150-400
    residues, seeds 5000+i), -fm 4, FirstPass (+ Refine with --refine); pairs sharded over the
    ranks by cost (minrms_b200/batch.py), no exchange.  A step = the whole batch."""
    import torch
    import torch.distributed as dist
    from minrms_b200 import Settings
    from minrms_b200.batch import all_vs_all, pair_cost
    from minrms_b200.synth import random_chain
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    chains = []
    for i in range(args.chains):
        rng = np.random.default_rng(5000 + i)
        chains.append(random_chain(rng, int(rng.integers(150, 401))))
    st = Settings(fm_fragment_size=4)
    lens = [c.shape[0] for c in chains]
    cells = sum(pair_cost(lens[i], lens[j], st) for i in range(len(lens)) for j in range(i + 1, len(lens)))
    times, ori = [], 0
    for it in range(args.warmup_batches + args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = all_vs_all(chains, st, rank, world, local, refine=args.refine, workers=args.workers)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0, float(sum(r.orientations for r in res.values())), float(len(res))],
                         dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tm[0]
        if it >= args.warmup_batches:
            times.append(float(t[0]))
        ori, npairs = int(t[1]), int(t[2])
    if rank == 0:
        sec = sum(times) / len(times)
        print(json.dumps({
            "metric": "orientations_per_sec", "value": ori / sec, "unit": "orientations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup_batches, "ms_per_step": 1e3 * sec, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeds 5000+i, minrms_b200/synth.py)",
            "config": {"workload": "all-vs-all of %d synthetic chains (%d..%d residues), -fm 4%s (config 5 shape)" % (
                len(lens), min(lens), max(lens), " + Refine" if args.refine else ", FirstPass only"),
                "config": 5, "pairs": npairs, "orientations": ori, "dp_cells": cells, "workers_per_gpu": args.workers,
                "parallelism": "pairs sharded over %d rank(s) by cost (LPT), no exchange" % world},
            "pairs_per_sec": npairs / sec, "dp_gcups": cells / sec / 1e9,
            "note": "end to end per batch: contexts, orientation generation, FirstPass%s, tables to the host" % (
                ", Refine" if args.refine else "")}))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--orientations", type=int, default=0, help="search only the first M orientations per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip full_run / nw_filter / strong-scaling extras")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 column-sweep, 2 tiled")
    ap.add_argument("--chains", type=int, default=8, help="config 5: number of chains")
    ap.add_argument("--workers", type=int, default=1, help="config 5: host threads (contexts) per GPU")
    ap.add_argument("--refine", action="store_true", help="config 5: Refine after FirstPass")
    ap.add_argument("--warmup-batches", type=int, default=1, help="config 5: untimed batches")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 5:
        return run_all_vs_all(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from minrms_b200 import OrientationGenerator, SearchOrCuda
    from minrms_b200.dist import gather_payload_device, partition, reduce_best_device, reduce_keys_device
    from minrms_b200.search import dp_cells_per_orientation

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(8, dtype=torch.int64, device=dev)
        dist.all_reduce(warm)            # create the communicator outside the timed regions
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = workload(args.config)
    A, B, st = w["A"], w["B"], w["settings"]
    m, nB = A.shape[0], B.shape[0]
    s = SearchOrCuda(A, B, st, device=local)
    if args.kernel:
        s.set_option("dp_kernel", args.kernel)
    if os.environ.get("MR_CHUNK"):
        s.set_option("chunk", int(os.environ["MR_CHUNK"]))
    if os.environ.get("MR_EXP_DP_ONLY"):
        s.set_option("exp_dp_only", 1)   # kernel experiments only: results are invalid
    og = OrientationGenerator(s)
    total = og.size()
    count = min(total, args.orientations) if args.orientations > 0 else total   # list searched per step
    cells_per = dp_cells_per_orientation(m, nB, s.n_min, s.n_max)
    strong = args.scaling == "strong" and world > 1
    if strong:
        first, stop = partition(count, rank, world)      # my share of ONE list, global ids
        id0, mine, job = first, stop - first, count
    else:
        first, id0, mine, job = 0, rank * count, count, count * world   # own copy, own ids
    pinned = torch.empty((count, 12), dtype=torch.float32).pin_memory()
    pinned.copy_(torch.from_numpy(og.orientations[:count]))
    host_orient = pinned.numpy()

    # ---------------- device-resident arm: `value` -----------------------------------
    s.upload_orientations(host_orient)

    # a step with the list resident: FirstPass + (N > 1) the cross-GPU MIN of the per-size keys in
    # device memory.  Like at N = 1, the winners' payload (transform + alignment) is fetched when
    # the table is wanted: once after the loop here, inside every step of the `e2e` arm.
    def step_resident():
        s.reset_best()
        s.search_resident(first, mine, id0)
        if world > 1:
            reduce_keys_device(s, dev)

    for _ in range(args.warmup):
        step_resident()
    if world > 1:
        gather_payload_device(s, dev)      # first use: row buffers, all_gather channels
    s.stats(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    search_ms = 0.0
    for _ in range(args.steps):
        s.event_record(0)
        s.reset_best()
        s.search_resident(first, mine, id0)
        s.event_record(1)
        search_ms += s.event_elapsed_ms(0, 1)        # CUDA events on the library's stream
        if world > 1:
            reduce_keys_device(s, dev)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    stats = s.stats()
    t = torch.tensor([search_ms / 1e3, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t[0]), float(t[1])
    payload_ms = 0.0
    if world == 1:
        s.fetch_best()
        table = s.best_table()
        step_s = dev_s / args.steps                 # device time (events), one GPU: nothing else in a step
    else:
        step_s = wall_s / args.steps                # barrier to barrier, max over ranks: search + key reduction
        barrier()
        p0 = time.perf_counter()
        table = gather_payload_device(s, dev)       # winners' payload from the owning ranks + all_gather + D2H
        barrier()
        tp = torch.tensor([time.perf_counter() - p0], dtype=torch.float64, device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        payload_ms = 1e3 * float(tp[0])
    red_ms = max(0.0, 1e3 * (wall_s - dev_s) / args.steps) if world > 1 else 0.0
    value = job / step_s

    # strong scaling: the merged table must be what one GPU finds on the whole list
    strong_check = None
    if strong and not args.no_extras:
        s.reset_best()
        s.search_resident(0, count, 0)
        s.fetch_best()
        solo = s.best_table()
        strong_check = bool(np.array_equal(solo["ssd"].view(np.uint32), table["ssd"].view(np.uint32)) and
                            np.array_equal(solo["o_id"], table["o_id"]) and
                            np.array_equal(solo["T"].view(np.uint32), table["T"].view(np.uint32)))

    # ---------------- end-to-end arm through the host-buffer C ABI: `e2e` -------------
    def step_e2e():
        s.reset_best()
        s.Search(host_orient, first, first + mine, first_id=id0)   # mr_search: H2D list, kernels, D2H table
        if world > 1:
            reduce_best_device(s, dev)

    e_steps = max(3, min(args.steps, 40))
    for _ in range(2):
        step_e2e()
    st_before = s.stats()
    barrier()
    e0 = time.perf_counter()
    for _ in range(e_steps):
        step_e2e()
    barrier()
    e_wall = time.perf_counter() - e0
    te = torch.tensor([e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    st_after = s.stats()
    e2e_value = job * e_steps / float(te[0])
    h2d = (st_after["h2d_bytes"] - st_before["h2d_bytes"]) // e_steps
    d2h = (st_after["d2h_bytes"] - st_before["d2h_bytes"]) // e_steps

    # ---------------- extras on one GPU: the whole run, and the NW filter ---------------
    full_run = nw_filter = alternates = None
    if world == 1 and not args.no_extras:
        reps = 3 if cells_per * total < 5e11 else 1
        evals = 0
        f0 = time.perf_counter()
        for rep in range(reps + (1 if reps > 1 else 0)):
            if rep == 1 and reps > 1:
                f0 = time.perf_counter()     # the first pass allocated this context's work space (cudaMalloc of GBs): warm from here
            s2 = SearchOrCuda(A, B, st, device=local)
            og2 = OrientationGenerator(s2)
            s2.Search(og2, 0, count)
            evals = s2.Refine() if np.any(s2.aSumSqdDists[:s2.R] < 1e26) else 0
            s2.close()
        f_s = (time.perf_counter() - f0) / reps
        full_run = {"value": count / f_s, "unit": "orientations/s", "ms": 1e3 * f_s, "refine_evaluations": int(evals),
                    "what": "context + orientation generation (GPU) + FirstPass + Refine + tables on the host, "
                            "through SearchOrCuda (the calls bin/minrms/main.cc:123-173 makes), warm process (work space of a previous context re-used)"}
        nw_filter = measure_nw_filter(local, clocks)
        if cells_per * total < 5e11:
            alternates = measure_alternates(SearchOrCuda, OrientationGenerator, A, B, st, local, count)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (3-D DP) ------------------------
    pk, pk_kind = peaks()
    f_mhz = (clocks or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    dp_ms = stats["dp_kernel_ms"]
    dp_gcups_kernel = stats["dp_cells"] / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else 0.0
    peak_gcups = sms * 128 * f_mhz * 1e6 / 3 / 1e9          # 1 FADD + 2 FMIN per cell
    peak_gcups_max = sms * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 3 / 1e9
    launches = max(stats["dp_kernel_launches"], 1)
    tb_bytes = 0.25 * stats["dp_cells"] / launches          # 2-bit traceback, algorithmic
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and w["cfg"] == 2:              # dram bytes from one ncu --set full capture
        tj = json.load(open(tpath))
        per_o = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["orientations_in_launch"]
        traffic = per_o * stats["orientations_dp"] / launches
    roofline = {"bound": "alu_issue", "achieved": dp_gcups_kernel, "peak": peak_gcups, "unit": "GCUPS",
                "frac": dp_gcups_kernel / peak_gcups if peak_gcups else None,
                "frac_at_max_clock": dp_gcups_kernel / peak_gcups_max,
                "peak_def": "%d SM x 128 lanes x f_SM / 3 ops per cell (1 FADD + 2 FMIN); f_SM = median SM clock sampled during the run (%.0f MHz)" % (sms, f_mhz),
                "kernel": "3-D DP (k_dp_*)", "launches": stats["dp_kernel_launches"],
                "avg_launch_ms": dp_ms / launches,
                "cells_per_launch": stats["dp_cells"] / launches,
                "traffic": traffic,
                "kernel_form": "k_dp_vp_* (single-pass problems): 1 FADD + 2 VIMNMX.U32 with predicate outputs + 2 predicated FADD per "
                               "cell, cubin patched after ptxas (minrms_b200/sass_patch.py); multi-pass problems: ping-pong form "
                               "(3 FADD + FMNMX3 + 2 SHF)",
                "measured_ceilings": {
                    "unit": "GCUPS",
                    "alu_pipe_two_ops_per_cell": sms * 64 * f_mhz * 1e6 / 2 / 1e9,
                    "issue_five_slots_per_cell": sms * 128 * f_mhz * 1e6 / 5 / 1e9,
                    "def": "every ALU-pipe instruction (VIMNMX, FMNMX, SHF, LOP3, MOV, ISETP) issues at 2 clk per warp "
                           "instruction per SM sub-partition, FADD at 1 (profiles/r02/sass_probe_vimnmx_predicates.log): the "
                           "min-with-predicate cell (2 ALU-pipe ops, 5 issue slots) is bounded by the issue rate, the round-1 "
                           "ping-pong cell (3 ALU-pipe ops, 6 slots) by the ALU pipe at 148 x 64 / 3 x f_SM"},
                "hbm": {"bound": "hbm", "achieved": tb_bytes / (dp_ms / launches * 1e-3) / 1e9 if dp_ms else 0.0,
                        "peak": pk["hbm_gbs"], "unit": "GB/s", "peak_kind": pk_kind,
                        "algorithmic_bytes_per_launch": tb_bytes,
                        "note": "2-bit traceback codes written once; the DP is issue-bound, not HBM-bound"},
                "share_of_step": {"dp": dp_ms / (dev_s * 1e3) if dev_s else None,
                                  "traceback_kabsch": stats["tb_kernel_ms"] / (dev_s * 1e3) if dev_s else None,
                                  "nw_filter": stats["nw_kernel_ms"] / (dev_s * 1e3) if dev_s else None}}

    cfgd = common_config(w, args, count)
    assert cfgd["dp_cells_per_orientation"] == int(cells_per), (cfgd["dp_cells_per_orientation"], cells_per)
    line = {"metric": "orientations_per_sec", "value": value, "unit": "orientations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
            "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": w["data"], "config": cfgd,
            "dp_gcups": value * cells_per * (stats["orientations_dp"] / max(stats["orientations"], 1)) / 1e9,
            "timed_region_s": wall_s,
            "search_ms_per_step": 1e3 * dev_s / args.steps,
            "reduction_ms": red_ms,
            "payload_ms_once_per_search": payload_ms,
            "gpu_launches": int(stats["kernel_launches"]),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "orientations/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e_steps},
            "roofline": roofline,
            "result_check": {"sizes_with_solution": int((table["ssd"] < 1e26).sum())}}
    if s.n_min <= 100 <= s.n_max and table["ssd"][100 - s.n_min] < 1e26:
        line["result_check"]["rmsd_100"] = float(np.sqrt(table["ssd"][100 - s.n_min] / 100.0))
    if strong_check is not None:
        line["result_check"]["merged_table_equals_one_gpu_search"] = strong_check
    if full_run:
        line["full_run"] = full_run
    if nw_filter:
        line["nw_filter"] = nw_filter
    if alternates:
        line["alternates"] = alternates

    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        if w["cfg"] in (1, 2):
            res = ref_bench(w, cores, ref_sample_size(w), 2)
            if res:
                n, tt, tot = res[-1]
                line["cpu_baseline"] = {"value": n / tt, "unit": "orientations/s", "cores": cores, "kind": "reference",
                                        "sample": "%d of the %d orientations (%d per process x %d processes), unmodified "
                                                  "reference Search(og,start,stop), shipped flags" % (n, tot, n // cores, cores)}
            else:
                line["cpu_baseline"] = None
        else:
            # one orientation costs minutes of a core: no bounded sample; the reference's measured cell rate instead
            per_core = CPU_CELLS_PER_S_PER_CORE / cells_per
            line["cpu_baseline"] = {"value": per_core * cores, "unit": "orientations/s", "cores": cores, "kind": "reference",
                                    "sample": "EXTRAPOLATED, not run: %.2e cells per orientation / 0.22 GCUPS per core (Solve_IncrN of "
                                              "the reference measured on config 2, SURVEY.md 6) x %d cores; DP only, the NW filter "
                                              "not counted" % (cells_per, cores)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_alternates(SearchOrCuda, OrientationGenerator, A, B, st, local, count):
    """The FirstPass variants behind -alt-o / -alt-r on the bench workload, host buffers, wall clock
    around the C ABI call (best of 5): mr_search, mr_search_scores (every compared score returned),
    mr_search_dissimilar with one lower rank (common-pairs test in the traceback kernel).  Extra
    information only: a failure here is reported in the block, never in the headline numbers."""
    try:
        s = SearchOrCuda(A, B, st, device=local)
        orient = np.ascontiguousarray(OrientationGenerator(s).orientations[:count])
        R, N = max(s.R, 1), s.N

        def fresh():
            s.load_table(dict(ssd=np.full(s.R, np.float32(1e26), np.float32), T=np.zeros((s.R, 12), np.float32),
                              pairs=np.full((R, N, 2), -1, np.int16)[:s.R], o_id=np.full(s.R, -1, np.int64)))

        def timed(f):
            best = None
            for _ in range(5):
                fresh()
                t0 = time.perf_counter()
                f()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            return 1e3 * best

        ms_search = timed(lambda: s.Search(orient))
        rank0 = s.best_table()
        ms_scores = timed(lambda: s.SearchScores(orient))
        same = bool(np.array_equal(s.best_table()["ssd"].view(np.uint32), rank0["ssd"].view(np.uint32)))
        ms_dis = timed(lambda: s.SearchDissimilar(orient, [rank0], max(1, N // 3)))
        found = int((s.best_table()["ssd"] < 1e26).sum())
        s.close()
        return {"search_ms": ms_search, "search_scores_ms": ms_scores, "search_dissimilar_1_lower_ms": ms_dis,
                "scores_bytes": int(4 * orient.shape[0] * s.R), "scores_table_equals_search": same,
                "max_pairs_common": max(1, N // 3), "second_rank_sizes_with_solution": found,
                "what": "mr_search / mr_search_scores / mr_search_dissimilar over the whole list, host buffers, best of 5"}
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def measure_nw_filter(local, clocks):
    """NW orientation filter (DynNW, lib/dyn/dynNW.cc:243-475) of config 1 on the test pair:
    18 532 orientations x 19 129 banded cells, timed alone over repeated launches"""
    from minrms_b200 import OrientationGenerator, SearchOrCuda
    import torch
    w = workload(1)
    s = SearchOrCuda(w["A"], w["B"], w["settings"], device=local)
    og = OrientationGenerator(s)
    s.upload_orientations(og.orientations)
    for _ in range(3):
        s.reset_best(); s.search_resident(0, og.size(), 0)
    s.stats(reset=True)
    reps = 20
    for _ in range(reps):
        s.reset_best(); s.search_resident(0, og.size(), 0)
    stt = s.stats()
    s.close()
    pk, _ = peaks()
    f_mhz = (clocks or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    gcups = stt["nw_cells"] / (stt["nw_kernel_ms"] * 1e-3) / 1e9 if stt["nw_kernel_ms"] > 0 else 0.0
    peak = sms * 128 * f_mhz * 1e6 / 5 / 1e9
    return {"workload": w["name"], "orientations": og.size(), "survivors": int(stt["orientations_dp"] // reps),
            "cells_per_orientation": int(stt["nw_cells"] // max(stt["orientations"], 1)),
            "avg_launch_ms": stt["nw_kernel_ms"] / reps,
            "roofline": {"bound": "alu_issue", "achieved": gcups, "peak": peak, "unit": "GCUPS", "frac": gcups / peak,
                         "peak_def": "%d SM x 128 lanes x f_SM / 5 ops per NW cell: 3 FADD (two gap steps + the match cost) and "
                                     "2 FMIN; the distance build (8 FP32 ops per cell, no FMA contraction), the strict-< tie "
                                     "order and the carried match count come on top" % sms}}


if __name__ == "__main__":
    main()
