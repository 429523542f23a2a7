#!/usr/bin/env python3
"""Rebuild the reference's test pair (tests/test_typical_usage.sh:5-8 downloads 4hhb/1a6m,
which needs a network) from the one in-repo fixture that embeds their CA coordinates:
/root/reference/tests/align_server_output/1a6m+4hhbA.txt  (SURVEY.md section 8c).

  block "IXYZ x_oldk, str. 1 original"  (151 CA)            -> myoglobin.pdb
  block "IXYZ x_new, str. 2 original", first 141 CA (chain A) -> hemoglobin.pdb

Writes CA-only fixed-column PDB files into tests/golden/.  Run once in the build
container (needs /root/reference); the outputs are committed.
"""
import os, sys

REF = os.environ.get("REF", "/root/reference")
SRC = os.path.join(REF, "tests/align_server_output/1a6m+4hhbA.txt")
HERE = os.path.dirname(os.path.abspath(__file__))


def blocks(path):
    out, cur, name = {}, None, None
    for line in open(path):
        if line.startswith("IXYZ DONE"):
            out[name] = cur
            cur = None
        elif line.startswith("IXYZ "):
            name = line[5:].strip()
            cur = []
        elif cur is not None and line.startswith("ATOM"):
            cur.append(line.rstrip("\n"))
    return out


def write_pdb(path, atoms):
    with open(path, "w") as f:
        for k, a in enumerate(atoms, 1):
            t = a.split()   # ATOM serial CA RES chain seq x y z  (not fixed-column in the fixture)
            resname, seq = t[3], int(t[5])
            x, y, z = float(t[6]), float(t[7]), float(t[8])
            f.write("ATOM  %5d  %-3s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n"
                    % (k, "CA", resname, "A", seq, x, y, z, 1.00, 0.00))
        f.write("END\n")


def main():
    b = blocks(SRC)
    myo = [v for k, v in b.items() if k.startswith("x_oldk, str. 1 original")][0]
    hemo = [v for k, v in b.items() if k.startswith("x_new, str. 2 original")][0]
    assert len(myo) == 151, len(myo)
    assert len(hemo) >= 141, len(hemo)
    write_pdb(os.path.join(HERE, "myoglobin.pdb"), myo)
    write_pdb(os.path.join(HERE, "hemoglobin.pdb"), hemo[:141])
    print("wrote myoglobin.pdb (151 CA), hemoglobin.pdb (141 CA) into", HERE)


if __name__ == "__main__":
    sys.exit(main())
