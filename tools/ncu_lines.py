"""Attribute the executed instructions / stall samples of an ncu report to CUDA source lines.

usage: python tools/ncu_lines.py report.ncu-rep 'sssp_nearfar_kernelILi256E' [top_n]
(joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` of the in-tree cubin by instruction order)"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(root, "streets4mpi_b200", "libs4mpi.so")], cwd=tmp,
                      stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], text=True).split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//") and ".text." in l and kern in l][0]
cur, inst = None, []
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        inst.append(cur)
rows = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True,
                                               stderr=subprocess.DEVNULL).split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ie, ia = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
src = open(os.path.join(root, "streets4mpi_b200", "csrc", "s4_kernels.cuh")).read().split("\n")
agg, st = collections.Counter(), collections.Counter()
for k in range(min(len(inst), len(data))):
    agg[inst[k]] += int(data[k][ie])
    st[inst[k]] += int(data[k][ia])
tot, tots = sum(agg.values()), sum(st.values())
print("SASS instructions: %d (ncu %d); executed warp instructions: %d" % (len(inst), len(data), tot))
for ln, c in agg.most_common(top_n):
    print("%5.1f%% inst %5.1f%% stall  L%-4s %s" % (100 * c / tot, 100 * st[ln] / tots, ln, src[ln - 1].strip()[:100] if ln else ""))
