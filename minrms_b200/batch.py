"""All-vs-all batch driver (BASELINE config 5: every unordered pair of a
set of chains).  The reference has no batch mode -- its man page calls all-vs-all searches
infeasible -- but every pair is an independent `minrms A B` run (bin/minrms/main.cc:129-173:
OrientationGenerator -> Search(og, 0, og.size()) -> Refine()), so the pairs shard across GPUs
with no exchange at all: each rank (one process per GPU) takes a cost-balanced share of the
pair list and runs the per-pair search on its own device through the C ABI.

    pairs  = schedule(lengths, world)[rank]            # (i, j), i < j; chain i is the fixed one
    result = all_vs_all(chains, settings, rank, world) # {(i, j): PairResult}

Nothing here computes: orientation generation, filter, DP, traceback, Kabsch, selection and
Refine all run in the CUDA library (minrms_b200/csrc)."""
import glob
import os
from dataclasses import dataclass, replace

import numpy as np

from .search import (OrientationGenerator, SearchOrCuda, Settings, dp_cells_per_orientation,
                     interpret_num_matches, load_ca_pdb, NoOrientationsError, UPPER_BOUND)


def pair_cost(m, n, settings):
    """DP cell updates of one pair: orientations x cells per orientation
    (the filter pass rate is unknown beforehand and ignored)."""
    L = settings.fm_fragment_size
    N = min(m, n)
    n_max = settings.n_max or N
    n_min = max(1, settings.n_min)
    if L > N or n_min > n_max:
        return 0
    return max(0, m - L + 1) * max(0, n - L + 1) * dp_cells_per_orientation(m, n, n_min, n_max)


def schedule(lengths, world, settings=None):
    """Cost-balanced partition of all unordered pairs over `world` ranks: pairs sorted by
    decreasing cost, each given to the least loaded rank (LPT).  Deterministic; every pair
    appears in exactly one rank's list, largest first within a rank."""
    settings = settings or Settings()
    pairs = [(i, j) for i in range(len(lengths)) for j in range(i + 1, len(lengths))]
    cost = {p: pair_cost(lengths[p[0]], lengths[p[1]], settings) for p in pairs}
    pairs.sort(key=lambda p: (-cost[p], p))
    load = [0] * world
    out = [[] for _ in range(world)]
    for p in pairs:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(p)
        load[r] += cost[p]
    return out


@dataclass
class PairResult:
    """per-size table of one pair, as SearchOrDyn3d leaves it (slot r = n - n_min)."""
    n_min: int
    nb: int
    ssd: np.ndarray        # [R] float32, 1e26 = no solution
    T: np.ndarray          # [R, 12] float32
    pairs: np.ndarray      # [R, 2N] int16 (first 2n entries of row r = the n matched index pairs)
    orientations: int
    orientations_dp: int
    refine_evals: int

    def rmsd(self, n):
        return float(np.sqrt(self.ssd[n - self.n_min] / np.float32(n * self.nb)))

    def sizes(self):
        return [self.n_min + r for r in range(self.ssd.shape[0]) if self.ssd[r] < UPPER_BOUND]


def pair_settings(settings, m, n, min_n=None, max_n=None, of_min=None):
    """Per-pair copy of `settings` with the reference's fractional arguments resolved for
    this pair (InterpretNumMatches, bin/minrms/minrms_parser.cc:1288-1327): min_n / max_n /
    of_min given as fractions (< 1) scale with min(m, n)."""
    s = replace(settings)
    if min_n is not None:
        s.n_min = max(1, interpret_num_matches(min_n, m, n))
    if max_n is not None:
        s.n_max = interpret_num_matches(max_n, m, n)
    if of_min is not None:
        s.orientation_filter_nw_min_num_matches = interpret_num_matches(of_min, m, n)
    return s


def search_pair(A, B, settings, device=0, refine=True, budget_mb=None):
    """One `minrms A B` run: FirstPass over every orientation, then Refine.
    budget_mb: cap of the chunk work space of this context (several contexts share a GPU)."""
    s = SearchOrCuda(A, B, settings, device=device)
    try:
        if budget_mb:
            s.set_option("budget_mb", budget_mb)
        og = OrientationGenerator(s)
        s.Search(og, 0, og.size())
        evals = 0
        if refine and np.any(s.aSumSqdDists[:s.R] < UPPER_BOUND):
            evals = s.Refine()
        st = s.stats()
        t = s.best_table()
        return PairResult(s.n_min, settings.nb, t["ssd"], t["T"], t["pairs"], int(st["orientations"]),
                          int(st["orientations_dp"]), evals)
    finally:
        s.close()


def all_vs_all(chains, settings=None, rank=0, world=1, device=0, refine=True, min_n=None,
               max_n=None, of_min=None, on_result=None, workers=1):
    """Search this rank's share of all unordered pairs of `chains` (list of float32 [len*nb, 3]
    arrays).  Pairs too short for the fragment size are skipped like the reference refuses them
    (lib/or_gen/or_gen.cc:548-580).  Returns {(i, j): PairResult}; `on_result(i, j, result)` is
    called as each pair finishes (e.g. to write it out).

    workers > 1: that many host threads take pairs from the rank's list, each with its own
    context (own stream and work space) on the same GPU, so that the set-up, the latency-bound
    Refine batches and the copies of one pair overlap the DP kernels of another.  The result
    of a pair does not depend on which thread ran it."""
    settings = settings or Settings()
    nb = settings.nb
    lengths = [c.shape[0] // nb for c in chains]
    mine = schedule(lengths, world, settings)[rank]
    out = {}
    # several contexts on one GPU share its memory: 140 GB of chunk work space between them
    budget_mb = None if workers <= 1 else min(32768, 140000 // workers)

    def run(p):
        i, j = p
        ps = pair_settings(settings, lengths[i], lengths[j], min_n, max_n, of_min)
        try:
            return p, search_pair(chains[i], chains[j], ps, device=device, refine=refine, budget_mb=budget_mb)
        except NoOrientationsError:
            return p, None     # the reference refuses such a pair; every other failure propagates

    if workers <= 1:
        done = map(run, mine)
    else:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=workers)
        done = pool.map(run, mine)     # results in list order, pairs run `workers` at a time
    for (i, j), res in done:
        if res is None:
            continue
        out[(i, j)] = res
        if on_result is not None:
            on_result(i, j, res)
    if workers > 1:
        pool.shutdown()
    return out


def main(argv=None):
    """python -m minrms_b200.batch DIR [--out results.npz] [--fm L] [--minN x] [--no-refine]
    [--workers k] [--msf-dir OUT]
    One rank per GPU under torchrun (RANK / WORLD_SIZE / LOCAL_RANK); each rank writes its own
    file (results.rank<k>.npz) -- there is nothing to reduce.  --msf-dir: also write, per pair,
    OUT/<a>__<b>/align<n>.msf in the reference's format (minrms_b200/msf.py)."""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("dir")
    ap.add_argument("--out", default="all_vs_all.npz")
    ap.add_argument("--fm", type=int, default=4)
    ap.add_argument("--minN", type=float, default=None)
    ap.add_argument("--no-refine", action="store_true")
    ap.add_argument("--workers", type=int, default=None,
                    help="host threads (contexts) per GPU; default 4 with Refine, 1 without")
    ap.add_argument("--msf-dir", default=None)
    a = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    device = int(os.environ.get("LOCAL_RANK", 0))
    files = sorted(glob.glob(os.path.join(a.dir, "*.pdb")))
    chains = [load_ca_pdb(f) for f in files]
    names = [os.path.basename(f) for f in files]
    seqs = None
    if a.msf_dir:
        from .msf import pdb_sequence, write_msf_files
        seqs = [pdb_sequence(f) for f in files]
        for f, c, q in zip(files, chains, seqs):   # the search and the sequence must index the same residues
            if c.shape[0] != len(q):
                raise SystemExit("%s: %d residues but %d with a CA atom; MSF output needs the index map "
                                 "(write_msf_files(index_a=..., index_b=...))" % (f, len(q), c.shape[0]))

    def report(i, j, r):
        sizes = r.sizes()
        print("%s %s: %d sizes, rmsd(N=%d) %.4f" % (names[i], names[j], len(sizes), sizes[-1] if sizes else 0,
                                                  r.rmsd(sizes[-1]) if sizes else 0.0), flush=True)
        if seqs is not None:
            d = os.path.join(a.msf_dir, "%s__%s" % (os.path.splitext(names[i])[0], os.path.splitext(names[j])[0]))
            os.makedirs(d, exist_ok=True)
            write_msf_files(d, r, seqs[i], seqs[j], names[i], names[j])

    workers = a.workers if a.workers else (1 if a.no_refine else 4)
    res = all_vs_all(chains, Settings(fm_fragment_size=a.fm), rank, world, device,
                     refine=not a.no_refine, min_n=a.minN, workers=workers, on_result=report)
    out = a.out if world == 1 else a.out.replace(".npz", ".rank%d.npz" % rank)
    np.savez_compressed(out, files=np.array(files), **{
        "%d_%d_%s" % (i, j, k): getattr(r, k) for (i, j), r in res.items() for k in ("ssd", "T", "pairs")})
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
