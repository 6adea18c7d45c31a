"""Diagnostic: do the same groups of trees take the same time in a whole day and in one GPU's share of it?"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from streets4mpi_b200 import engine, synth  # noqa: E402
from streets4mpi_b200.settings import settings  # noqa: E402

dev = engine.default_device()
arr = synth.grid_arrays(1024, 1024, seed=1)
V = len(arr["node_ids"])
cats = synth.node_categories(arr["node_ids"], seed=1)
flat = synth.network_from(arr).flat()
curve = flat.locality_rank()
goals = cats["goals"]
along = goals[np.argsort(curve[goals], kind="stable")]
phys = engine.make_physics(settings)
dev.set_option("root_mode", 1)
dev.set_option("group_log", 1)
for kv in filter(None, (sys.argv[1] if len(sys.argv) > 1 else "").split(",")):
    k, v = kv.split("=")
    dev.set_option(k, float(v))
g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"], node_rank=curve)
rng = np.random.default_rng(0)


def run(roots, name):
    o = rng.choice(arr["node_ids"], size=len(roots)).astype(np.int32)
    sim = engine.DeviceSim(g, o, roots.astype(np.int32), 0.4, phys)
    import time
    for _ in range(2):
        sim.reset_counters()
        dev.sync()
        t0 = time.perf_counter()
        sim.step()
        dev.sync()
        wall = (time.perf_counter() - t0) * 1e3
        c = sim.counters()
        sim.commit_total()
    slots, K = sim.root_slots()
    gl = g.group_log().astype(np.float64)
    groups = slots.reshape(-1, K)
    ms = gl[:, 0] * 1024 / 1.965e6
    info = {}
    for i, row in enumerate(groups):
        key = tuple(sorted(int(x) for x in row if x >= 0))
        info[key] = (ms[i], gl[i, 1], gl[i, 2] / V, i)
    print(json.dumps({"run": name, "roots": len(roots), "groups": len(groups), "wall_ms": round(wall, 1), "ms_sssp": round(c["ms_sssp"], 1), "launches": c["sssp_launches"],
                      "group_ms_pct": [round(float(x), 1) for x in np.percentile(ms, [0, 10, 50, 90, 99, 100])],
                      "rounds_pct": [float(x) for x in np.percentile(gl[:, 1], [0, 10, 50, 90, 99, 100])]}), flush=True)
    sim.close()
    return info


full = run(along, "full day")
n8, n4 = len(along) // 8, len(along) // 4
for name, part in (("first half", along[: len(along) // 2]), ("2nd quarter", along[n4: 2 * n4]), ("4th eighth", along[3 * n8: 4 * n8]),
                   ("1st eighth", along[:n8])):
    sub = run(part, name)
    common = [k for k in sub if k in full]
    if common:
        a = np.array([full[k][:3] for k in common])
        b = np.array([sub[k][:3] for k in common])
        print(json.dumps({"run": name, "identical_groups": len(common),
                          "ms_full_vs_part_median": [round(float(np.median(a[:, 0])), 1), round(float(np.median(b[:, 0])), 1)],
                          "rounds_full_vs_part_median": [float(np.median(a[:, 1])), float(np.median(b[:, 1]))],
                          "exp_per_node_full_vs_part_median": [round(float(np.median(a[:, 2])), 2), round(float(np.median(b[:, 2])), 2)],
                          "first_examples": [[full[k][3], sub[k][3], round(full[k][0], 1), round(sub[k][0], 1), full[k][1], sub[k][1]] for k in common[:8]]}), flush=True)
