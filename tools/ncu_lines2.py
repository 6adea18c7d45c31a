"""Attribute executed warp instructions / stall samples of an ncu report to CUDA source lines of any file.
usage: python tools/ncu_lines2.py report.ncu-rep 'mangled-kernel-substring' [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(root, "streets4mpi_b200", "libs4mpi.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], text=True).split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//") and ".text." in l and kern in l][0]
cur, inst = None, []
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        inst.append((cur, l.strip()))
rows = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], text=True, stderr=subprocess.DEVNULL).split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ie, ia = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
print("instructions: nvdisasm %d, ncu %d" % (len(inst), len(data)))
agg, st = collections.Counter(), collections.Counter()
for k in range(min(len(inst), len(data))):
    agg[inst[k][0]] += int(data[k][ie]); st[inst[k][0]] += int(data[k][ia])
tot, tots = sum(agg.values()), sum(st.values())
srcs = {}
for (f, ln), n in agg.most_common(top_n):
    if f not in srcs:
        srcs[f] = open(os.path.join(root, "streets4mpi_b200", "csrc", f)).read().split("\n") if os.path.exists(os.path.join(root, "streets4mpi_b200", "csrc", f)) else []
    text = srcs[f][ln - 1].strip()[:110] if srcs[f] and ln - 1 < len(srcs[f]) else ""
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * n / tot, 100.0 * st[(f, ln)] / max(1, tots), f, ln, text))
