#!/usr/bin/env python3
"""bench.py -- MINRMS per-orientation search throughput on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the reference's own test pair -- hemoglobin chain A
(141 CA) vs myoglobin (151 CA), tests/test_typical_usage.sh -- searched UNCONSTRAINED:
`minrms -fm 4 -HS -ir` = 20 424 orientations x 1 044 481 DP cells, all 140 alignment sizes
refined by Kabsch for every orientation.  One "step" = one complete FirstPass over the
orientation list of this rank.  With N GPUs every rank gets its own copy of the list under
distinct orientation ids (weak scaling: per-GPU work fixed), followed by the NCCL
min-reduction of the per-size best records.

Prints ONE JSON line (rank 0).  `value` = orientations/s with the orientation list already
resident in HBM; `e2e` = the same through the host-buffer C-ABI call (mr_search: H2D of the
list from pinned memory + D2H of the best table inside the timed region).
`--impl reference` times the UNMODIFIED reference (oracle/_ref/fast, built by
oracle/build_ref.sh) on the host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_ARGS = ["-fm", "4", "-HS", "-ir", "-R"]   # config 2, FirstPass only (the timed path)
WORKLOAD = "hemoglobin(4hhb:A,141)xmyoglobin(1a6m,151) unconstrained: minrms -fm 4 -HS -ir (config 2)"
M, NB, NMIN, NMAX, FRAG = 141, 151, 1, 141, 4


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


def ref_bench(procs, per_proc, repeat):
    """the reference's own Search(og,start,stop) on `procs` host processes (oracle/ref_bench.cc)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "fast", "ref_bench")
    if not os.path.exists(exe):
        return None
    out = subprocess.run([exe, str(procs), str(per_proc), str(repeat)] + REF_ARGS +
                         ["hemoglobin.pdb,*.A", "myoglobin.pdb"], cwd=GOLDEN, capture_output=True, text=True)
    res = []
    for line in out.stdout.splitlines():
        if line.startswith("ref_bench"):
            kv = dict(x.split("=") for x in line.split()[1:])
            if int(kv["failed"]) == 0:
                res.append((int(kv["orientations"]), float(kv["seconds"])))
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_proc = 160                       # ~1 s per process per step at ~160 orientations/s/core
    cells = 1044481
    res = ref_bench(cores, per_proc, args.warmup + args.steps)
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast/ref_bench not built"}))
        return
    res = res[args.warmup:]
    n = sum(r[0] for r in res)
    t = sum(r[1] for r in res)
    v = n / t
    sample = "%d orientations/step (%d per process x %d processes) of the 20424, Search(og,start,stop) of the unmodified reference, shipped flags -O3 -ffast-math" % (res[0][0], per_proc, cores)
    line = {"metric": "orientations_per_sec", "value": v, "unit": "orientations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(res),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "fixture", "impl": "reference", "dp_gcups": v * cells / 1e9,
            "config": {"workload": WORKLOAD, "orientations_per_step": res[0][0]},
            "cpu_baseline": {"value": v, "unit": "orientations/s", "cores": cores, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": v, "unit": "orientations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 column-sweep, 2 tiled")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from minrms_b200 import OrientationGenerator, SearchOrCuda, Settings, load_ca_pdb
    from minrms_b200.dist import allreduce_best

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

    A = load_ca_pdb(os.path.join(GOLDEN, "hemoglobin.pdb"))
    B = load_ca_pdb(os.path.join(GOLDEN, "myoglobin.pdb"))
    st = Settings(n_min=NMIN, n_max=NMAX, fm_fragment_size=FRAG)
    s = SearchOrCuda(A, B, st, device=local)
    if args.kernel:
        s.set_option("dp_kernel", args.kernel)
    if os.environ.get("MR_CHUNK"):
        s.set_option("chunk", int(os.environ["MR_CHUNK"]))
    if os.environ.get("MR_EXP_DP_ONLY"):
        s.set_option("exp_dp_only", 1)   # kernel experiments only: results are invalid
    og = OrientationGenerator(s)
    count = og.size()
    cells_per = s.lib.mr_dp_cells_per_orientation(M, NB, NMIN, NMAX)
    id0 = rank * count                      # this rank's ids (weak scaling: own copy of the list)
    pinned = torch.empty((count, 12), dtype=torch.float32).pin_memory()
    pinned.copy_(torch.from_numpy(og.orientations))
    host_orient = pinned.numpy()

    # ---------------- device-resident arm: `value` -----------------------------------
    s.upload_orientations(host_orient)

    def step_resident():
        s.reset_best()
        s.search_resident(0, count, id0)

    for _ in range(args.warmup):
        step_resident()
    s.stats(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    s.event_record(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
    s.event_record(1)
    dev_ms = s.event_elapsed_ms(0, 1)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    stats = s.stats()
    t = torch.tensor([dev_ms / 1e3, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t[0]), float(t[1])
    # the (tiny) cross-GPU reduction of the per-size records, timed separately
    s.fetch_best()
    table = s.best_table()
    red_ms = 0.0
    if world > 1:
        allreduce_best(table, device=dev)    # first use sets up the NCCL channels
        barrier()
        r0 = time.perf_counter()
        table = allreduce_best(table, device=dev)
        barrier()
        red_ms = 1e3 * (time.perf_counter() - r0)
    # N > 1: every search ends with the cross-GPU reduction of the per-size records, so its
    # time (max over ranks, measured above) is charged to every step of `value`
    step_s = dev_s / args.steps + red_ms * 1e-3
    value = count * world / step_s

    # ---------------- end-to-end arm through the host-buffer C ABI: `e2e` -------------
    for _ in range(2):
        s.reset_best(); s.Search(host_orient)
    st_before = s.stats()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        s.reset_best()
        s.Search(host_orient)               # mr_search: H2D list, kernels, D2H table
        if world > 1:
            allreduce_best(s.best_table(), device=dev)
    barrier()
    e_wall = time.perf_counter() - e0
    te = torch.tensor([e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    st_after = s.stats()
    e2e_value = count * world * args.steps / float(te[0])
    h2d = (st_after["h2d_bytes"] - st_before["h2d_bytes"]) // args.steps
    d2h = (st_after["d2h_bytes"] - st_before["d2h_bytes"]) // args.steps

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
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):                                # dram bytes from one ncu --set full capture
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
                "alu_pipe_ceiling": {"peak": sms * 64 * f_mhz * 1e6 / 2 / 1e9, "unit": "GCUPS",
                                     "def": "2 FMIN per cell on the half-rate ALU pipe (64 lanes/clk/SM)",
                                     "frac": dp_gcups_kernel / (sms * 64 * f_mhz * 1e6 / 2 / 1e9)},
                "hbm": {"bound": "hbm", "achieved": tb_bytes / (dp_ms / launches * 1e-3) / 1e9 if dp_ms else 0.0,
                        "peak": pk["hbm_gbs"], "unit": "GB/s", "peak_kind": pk_kind,
                        "algorithmic_bytes_per_launch": tb_bytes,
                        "note": "2-bit traceback codes written once; the DP is issue-bound, not HBM-bound"},
                "share_of_step": {"dp": dp_ms / (dev_s * 1e3) if dev_s else None,
                                  "traceback_kabsch": stats["tb_kernel_ms"] / (dev_s * 1e3) if dev_s else None}}

    line = {"metric": "orientations_per_sec", "value": value, "unit": "orientations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "fixture (CA traces of the reference's test pair, tests/golden/*.pdb)",
            "config": {"workload": WORKLOAD, "orientations_per_gpu_per_step": count,
                       "dp_cells_per_orientation": int(cells_per), "alignment_sizes": NMAX - 1,
                       "l2": "no explicit flush: each step writes and re-reads >= %.1f GB of traceback tables per GPU (>> 126 MB L2)" % (0.25 * cells_per * count / 1e9),
                       "parallelism": "orientations partitioned, %d rank(s), NCCL min-reduction of per-size keys (reduction_ms is charged to every step of `value`)" % world},
            "dp_gcups": value * cells_per / 1e9,
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "reduction_ms": red_ms,
            "gpu_launches": int(stats["kernel_launches"]),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "orientations/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "roofline": roofline,
            "result_check": {"sizes_with_solution": int((table["ssd"] < 1e26).sum()),
                             "rmsd_100": float(np.sqrt(table["ssd"][100 - NMIN] / 100.0))}}

    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        res = ref_bench(cores, 160, 2)
        if res:
            n, tt = res[-1]
            line["cpu_baseline"] = {"value": n / tt, "unit": "orientations/s", "cores": cores, "kind": "reference",
                                    "sample": "%d of the 20424 orientations (160 per process x %d processes), unmodified reference Search(og,start,stop), shipped flags" % (n, cores)}
        else:
            line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
