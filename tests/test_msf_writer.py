"""minrms_b200/msf.py against the files the UNMODIFIED reference binary wrote: for the
reference's own final tables (tests/golden/ref_<cfg>.npz, strict build, after Refine) every
align<n>.msf must hash to the digest in tests/golden/msf_digests.json (time-stamp line masked,
tests/golden/make_msf_digests.py).  CPU only: the tables are the committed ones."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
from make_msf_digests import MOLS, PAIR, digest_dir  # noqa: E402

from minrms_b200.msf import msf_text, pdb_sequence, write_msf_files  # noqa: E402


@pytest.mark.parametrize("cfg", ["c1", "c2", "long", "bb"])
def test_msf_files_equal_the_reference_binary_s(cfg, tmp_path):
    g = np.load(os.path.join(GOLD, "ref_%s.npz" % cfg))
    want = json.load(open(os.path.join(GOLD, "msf_digests.json")))[cfg]
    mols = [f.split(",")[0] for f in MOLS.get(cfg, PAIR)]
    seqs = [pdb_sequence(os.path.join(GOLD, f)) for f in mols]
    assert len(seqs[0]) == int(g["m"]) and len(seqs[1]) == int(g["nB"])
    table = dict(ssd=g["strict_refine_ssd"], T=g["strict_refine_T"], pairs=g["strict_refine_pairs"],
                 n_min=int(g["n_min"]), nb=int(g["nb"]))
    files = write_msf_files(str(tmp_path), table, seqs[0], seqs[1], mols[0], mols[1])
    got = digest_dir(str(tmp_path))
    want = {k: v for k, v in want.items() if k.endswith(".msf")}
    assert sorted(got) == sorted(want) == sorted(files)
    diff = [f for f in want if got[f] != want[f]]
    assert not diff, "files differ from the reference: %s" % diff[:5]


def test_msf_layout_small_case():
    txt = msf_text("align2.msf", 2, 0.5, np.arange(12), [[1, 0], [3, 2]], "ACDEF", "GHIK", "a.pdb", "bb.pdb", "bb.pdb",
                   when=__import__("time").struct_time((2026, 10, 19, 7, 5, 0, 0, 0, 0)))
    lines = txt.split("\n")
    assert lines[0] == "Chimera minimal RMSD structural alignment with 2 equivalences."
    assert lines[1] == "RMSD = 0.5" and lines[4] == "0 1 2 3"
    assert lines[8] == "align2.msf   MSF: 7  Type: ?  October 19, 2026  7:5  Check: 0 .."
    assert lines[10] == "Name: a.pdb   Len: 5   Check: 0   Weight: 1.00"
    assert lines[11] == "Name: bb.pdb  Len: 4   Check: 0   Weight: 1.00"
    # A: A C . D E F   B: . G H . I K -> unmatched A first, then unmatched B, then the pair
    # the end label is placed as if the line were full (5 group separators), like the reference
    assert lines[15] == "        1          7"
    assert lines[16].startswith("a.pdb   ") and lines[17].startswith("bb.pdb  ")
    a, b = lines[16][8:], lines[17][8:]
    assert a.replace(".", "") == "ACDEF" and b.replace(".", "") == "GHIK" and len(a) == len(b) == 7
    assert [k for k in range(7) if a[k] != "." and b[k] != "."] == [1, 4]
    with pytest.raises(ValueError):
        msf_text("x.msf", 2, 0.5, np.arange(12), [[3, 0], [1, 2]], "ACDEF", "GHIK", "a", "b", "b")
