"""Seeded synthetic chains for the benchmark configurations and the tests (SURVEY.md 8d:
CA-trace random walks with 3.8 A bonds, coordinates rounded to the 3 decimals of a PDB file),
and a fixed-column PDB writer so that the reference binary can read the same chains."""
import numpy as np


def random_chain(rng, n, nb=1):
    """seeded CA-trace random walk, 3.8 A bonds, coordinates rounded to 3 decimals"""
    pts = np.zeros((n, 3))
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    for k in range(1, n):
        d = d + 0.8 * rng.normal(size=3)
        d /= np.linalg.norm(d)
        pts[k] = pts[k - 1] + 3.8 * d
    if nb > 1:
        pts = np.repeat(pts, nb, axis=0) + rng.normal(scale=0.7, size=(n * nb, 3))
    return np.round(pts, 3).astype(np.float32)


def related_chain(rng, A, nB, nb=1, noise=1.0, return_idx=False):
    """a rigidly moved, noisy copy of A with indels, length nB residues"""
    m = A.shape[0] // nb
    idx = np.sort(rng.choice(m, size=min(m, nB), replace=False))
    base = A.reshape(m, nb, 3)[idx]
    if nB > base.shape[0]:
        extra = random_chain(rng, nB - base.shape[0], nb).reshape(-1, nb, 3) + base[-1, 0]
        base = np.concatenate([base, extra])
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    Rm = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                   [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                   [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
    out = base.reshape(-1, 3) @ Rm.T + rng.normal(scale=10, size=3) + rng.normal(scale=noise, size=(nB * nb, 3))
    out = np.round(out, 3).astype(np.float32)
    return (out, idx) if return_idx else out


def write_ca_pdb(path, coords, chain="A"):
    """CA-only PDB in the fixed columns the reference parses (lib/pdb++/read_format.i:21):
    one ALA residue per point"""
    coords = np.asarray(coords, np.float32).reshape(-1, 3)
    with open(path, "w") as f:
        for k, (x, y, z) in enumerate(coords):
            f.write("ATOM  %5d  %-3s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" %
                    ((k + 1) % 100000, "CA", "ALA", chain, (k + 1) % 10000, x, y, z, 1.0, 0.0))
        f.write("END\n")
