import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_c1():
    return _load("ref_c1.npz")


@pytest.fixture(scope="session")
def golden_c2():
    return _load("ref_c2.npz")


@pytest.fixture(scope="session")
def golden_mid():
    """230 x 250 synthetic pair, `-fm 4,0.6 -HS -ir -r` (887 orientations): run of the
    unmodified reference on a problem that needs the multi-pass DP kernel"""
    return _load("ref_mid.npz")


@pytest.fixture(scope="session")
def golden_long():
    """280 x 300 synthetic pair, `-fm 4,0.6 -of 8.0,0.33 -HS -minN 0.33 -max-rmsd 4.0 -ir -r`:
    the flags of config 1 on a pair whose second molecule is longer than 256 residues"""
    return _load("ref_long.npz")


@pytest.fixture(scope="session")
def golden_bb():
    """60 x 66 residues with four backbone atoms each (atoms_used.txt = N, C, CA, O), unconstrained"""
    return _load("ref_bb.npz")


@pytest.fixture(scope="session")
def golden_bb_of():
    """the same pair under `-of 8.0,0.33 -minN 0.33 -max-rmsd 3.0`"""
    return _load("ref_bb_of.npz")


@pytest.fixture(scope="session")
def golden_ir_off():
    """CA atoms of the 60 x 66 pair, `-IR -max-dist 7.0 -minN 0.25 -r`"""
    return _load("ref_ir_off.npz")


@pytest.fixture(scope="session")
def golden_iv():
    """CA atoms of the 60 x 66 pair, two `-fm-interval` pairs restricting the fragment starts"""
    return _load("ref_iv.npz")


@pytest.fixture(scope="session")
def golden_maxn():
    """CA atoms of the 60 x 66 pair, `-fm 5 -minN 10 -maxN 40`"""
    return _load("ref_maxn.npz")
