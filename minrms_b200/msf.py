"""align<n>.msf files for the all-vs-all driver (per-pair MSF writing).

The drop-in binary keeps the reference's own writer; the batch driver has no reference code
behind it, so this module restates the text format of SearchOrDyn3d::WriteMSF
(bin/minrms/search_or_dyn3d.cc:2848-2940) and PairwiseAlignment::ExportMSF
(lib/pair_alignment/pair_alignment.cc:1470-1920) with their default arguments (number lines
shown, no lower-case compression, first solution).  tests/test_msf_writer.py checks the files
byte for byte (time-stamp line masked) against the digests of what the unmodified reference
binary wrote.  Host-side formatting only: every number printed comes from the CUDA search."""
import time

import numpy as np

GAP_CHAR = "."          # PairwiseAlignment::GAP_CHAR, pair_alignment.cc:53
UNKNOWN_RESIDUE = "x"   # NameCodeLookup::UNKNOWN_RESIDUE, lib/biopolymer/residue_names.cc:41
# the built-in 3-letter -> 1-letter table (residue_names.cc:55-130; res_code_dict.txt extends it)
RES_CODES = {"GLY": "G", "ALA": "A", "SER": "S", "CYS": "C", "VAL": "V", "THR": "T", "ILE": "I",
             "PRO": "P", "MET": "M", "ASP": "D", "ASN": "N", "LEU": "L", "LYS": "K", "GLU": "E",
             "GLN": "Q", "ARG": "R", "HIS": "H", "PHE": "F", "TYR": "Y", "TRP": "W"}
MONTHS = ["January", "February", "March", "April", "May", "June", "July", "August", "September",
          "October", "November", "December"]


def pdb_sequence(path, chain=None, codes=None):
    """One-letter sequence of the ATOM residues of a PDB file in file order (Mol2Sequence,
    lib/biopolymer/mol2sequence.cc:47-80): one letter per residue, 'x' for unknown names."""
    codes = codes or RES_CODES
    seq, seen = [], set()
    with open(path) as f:
        for line in f:
            if not line.startswith("ATOM"):
                continue
            ch = line[21]
            if chain is not None and ch != chain:
                continue
            key = (ch, line[22:27])
            if key in seen:
                continue
            seen.add(key)
            seq.append(codes.get(line[17:20].strip(), UNKNOWN_RESIDUE))
    return "".join(seq)


def _g(x):
    """default ostream formatting of a float (6 significant digits)"""
    return "%g" % float(x)


def msf_text(filename, n, rmsd, T, pairs, seq_a, seq_b, label_a, label_b, moved_pdb, when=None):
    """Text of one align<n>.msf.  pairs: n x 2 residue indices into seq_a / seq_b (ascending);
    T: the 12 floats of the stored transform; rmsd as WriteMSF computes it,
    sqrt(ssd / (n * nb)) in float."""
    pairs = np.asarray(pairs).reshape(-1, 2)[:n]
    if np.any(np.diff(pairs[:, 0]) <= 0) or np.any(np.diff(pairs[:, 1]) <= 0):
        raise ValueError("Error writing MSF file: the alignment is not sequential (pair_alignment.cc:1483-1489)")
    T = np.asarray(T, np.float32).reshape(3, 4)
    out = ["Chimera minimal RMSD structural alignment with %d equivalences.\n" % n,
           "RMSD = %s\n" % _g(rmsd), "-----\n",
           "Transform Matrix to apply to structure: %s\n" % moved_pdb]
    for r in range(3):
        out.append(" ".join(_g(v) for v in T[r]) + "\n")
    out.append("\n")                                   # `comments << "\n"`
    la, lb = len(seq_a), len(seq_b)
    total = (la - n) + (lb - n) + n
    t = time.localtime() if when is None else when
    out.append("%s   MSF: %d  Type: ?  %s %d, %d  %d:%d  Check: 0 ..\n\n" % (
        filename, total, MONTHS[t.tm_mon - 1], t.tm_mday, t.tm_year, t.tm_hour, t.tm_min))
    label_size = max(len(label_a), len(label_b)) + 2
    pad_a, pad_b = " " * (label_size - len(label_a)), " " * (label_size - len(label_b))
    out.append("Name: %s%sLen: %d   Check: 0   Weight: 1.00\n" % (label_a, pad_a, la))
    out.append("Name: %s%sLen: %d   Check: 0   Weight: 1.00\n" % (label_b, pad_b, lb))
    out.append("\n//\n\n")
    # the two gapped rows: unmatched residues of A, then of B, then the matched pair
    row_a, row_b = [], []
    pi, pj = -1, -1
    for k in range(n + 1):
        ci, cj = (int(pairs[k, 0]), int(pairs[k, 1])) if k < n else (la, lb)
        row_a.append(seq_a[pi + 1:ci]); row_b.append(GAP_CHAR * (ci - pi - 1))
        row_a.append(GAP_CHAR * (cj - pj - 1)); row_b.append(seq_b[pj + 1:cj])
        if k < n:
            row_a.append(seq_a[ci]); row_b.append(seq_b[cj])
        pi, pj = ci, cj
    row_a, row_b = "".join(row_a), "".join(row_b)
    assert len(row_a) == len(row_b) == total
    per_line, per_group = 50, 10
    for c0 in range(0, total, per_line):
        nchars = min(per_line, total - c0)
        begin, end = str(c0 + 1), str(c0 + nchars)
        nspaces = (c0 + per_line + 1) // per_group - (c0 + 1) // per_group
        between = max(nchars + nspaces - (len(begin) + len(end)), 1)
        out.append(" " * label_size + begin + " " * between + end + "\n")
        for label, pad, row in ((label_a, pad_a, row_a), (label_b, pad_b, row_b)):
            line = [label, pad]
            for c in range(c0, c0 + nchars):
                line.append(row[c])
                if (c + 1) % per_group == 0:
                    line.append(" ")
            out.append("".join(line) + "\n")
        out.append("\n")
    return "".join(out)


def write_msf_files(directory, result, seq_a, seq_b, label_a, label_b, moved_pdb=None, index_a=None,
                    index_b=None, base="align"):
    """Write align<n>.msf for every size of a PairResult / best table that has a solution
    (the loop of SearchOrDyn3d::WriteMSF).  index_a / index_b map the residues the search saw
    to positions in seq_a / seq_b (TranslateAlignmentBetweenMolecules, search_or_dyn3d.cc:2912-2917);
    None = identity.  Returns the list of files written."""
    import os
    ssd = np.asarray(result.ssd if hasattr(result, "ssd") else result["ssd"])
    T = np.asarray(result.T if hasattr(result, "T") else result["T"])
    pairs = np.asarray(result.pairs if hasattr(result, "pairs") else result["pairs"])
    n_min = result.n_min if hasattr(result, "n_min") else result["n_min"]
    nb = result.nb if hasattr(result, "nb") else result.get("nb", 1)
    pairs = pairs.reshape(ssd.shape[0], -1, 2)
    written = []
    for r in range(ssd.shape[0]):
        if not ssd[r] < np.float32(1.0e26):
            continue
        n = n_min + r
        rmsd = np.sqrt(np.float32(ssd[r]) / np.float32(n * nb), dtype=np.float32)
        p = pairs[r, :n].astype(np.int64)
        if index_a is not None:
            p[:, 0] = np.asarray(index_a)[p[:, 0]]
        if index_b is not None:
            p[:, 1] = np.asarray(index_b)[p[:, 1]]
        name = "%s%d.msf" % (base, n)
        with open(os.path.join(directory, name), "w") as f:
            f.write(msf_text(name, n, rmsd, T[r], p, seq_a, seq_b, label_a, label_b, moved_pdb or label_b))
        written.append(name)
    return written
