"""Summarise `nvcc -Xptxas -v` output of the library: registers / spills per kernel (demangled)."""
import re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minrms_b200 import build as b
cmd = ["nvcc"] + b.NVCC_FLAGS + os.environ.get("MR_NVCC_EXTRA", "").split() + ["-Xptxas", "-v", "-o", os.environ.get("MR_BUILD_OUT", b.OUT), b.SRC]
out = subprocess.run(cmd, capture_output=True, text=True)
txt = out.stderr + out.stdout
if out.returncode: print(txt[-3000:]); sys.exit(1)
names = re.findall(r"Compiling entry function '(\S+)'", txt)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"Compiling entry function ", txt)[1:]
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for nm, blk in zip(dem, blocks):
    short = re.sub(r"\(.*", "", nm).replace("void ", "")
    if pat and pat not in short: continue
    regs = re.search(r"Used (\d+) registers", blk)
    sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    print("%-70s regs %3s  spill %s/%s" % (short[:70], regs.group(1) if regs else "?", sp.group(1), sp.group(2)))
