"""Turns the ncu CSV logs written by scripts/ncu_capture.sh into the tracked summaries under profiles/:
     profiles/<tag>_launch_list_step.csv    one row per launch: kernel, grid, block, device time (ns)
     profiles/<tag>_kernel_metrics_step.csv one row per kernel name: launches in the window + mean of every metric
   usage: python scripts/summarize_ncu.py <tag>      (reads gpurun_out/launches_<tag>.csv and gpurun_out/kernels_<tag>.csv)"""
import csv, os, re, sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def rows_of(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def one_step(rows):
    """keep the launches of ONE train step.  Steps repeat the same kernel sequence, so the window (longer than a step)
    is periodic: find the period S from the kernel names and keep S consecutive launches (any S consecutive launches
    are one step's worth of work)."""
    ids = sorted({int(r["ID"]) for r in rows})
    name_of = {}
    for r in rows:
        name_of.setdefault(int(r["ID"]), r["Kernel Name"])
    names = [name_of[i] for i in ids]
    n = len(names)
    S = None
    for cand in range(50, n):
        if all(names[i] == names[i + cand] for i in range(n - cand)) and n - cand >= 8:
            S = cand
            break
    if S is None:
        return rows
    keep = set(ids[:S])
    return [r for r in rows if int(r["ID"]) in keep]


launches = one_step(rows_of(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")))
with open(os.path.join(ROOT, "profiles", f"{tag}_launch_list_step.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum[ns]"])
    tot = 0.0
    for r in launches:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        w.writerow([r["ID"], short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], r["Metric Value"]])
        tot += float(r["Metric Value"].replace(",", ""))
print(f"launch list: {len(launches)} launches, {tot/1e6:.3f} ms (cold-cache, serialised)")

metrics = one_step(rows_of(os.path.join(ROOT, "gpurun_out", f"kernels_{tag}.csv")))
per = OrderedDict()
ids = defaultdict(set)
for r in metrics:
    k = short(r["Kernel Name"])
    ids[k].add(r["ID"])
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    per.setdefault(k, defaultdict(list))[r["Metric Name"]].append(v)
names = sorted({m for d in per.values() for m in d})
with open(os.path.join(ROOT, "profiles", f"{tag}_kernel_metrics_step.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches_in_window"] + names)
    for k, d in per.items():
        w.writerow([k, len(ids[k])] + [round(sum(d[m]) / len(d[m]), 3) if d.get(m) else "" for m in names])
print(f"kernel metrics: {len(per)} kernels")
