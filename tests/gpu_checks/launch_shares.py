"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py to (a) a compact launch list
(id, kernel, grid, block, duration_us) and (b) per-kernel sums for ONE step: the longest run of launches from one reset_kernel up to
the next (or the end of the list).

    python tests/gpu_checks/launch_shares.py gpurun_out/x_raw.csv profiles/x_launches.csv profiles/x_launch_shares.csv
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    name = name.replace("(anonymous namespace)", "<unnamed>")
    return name.strip()


def main(raw, out_list, out_shares):
    rows = []
    with open(raw) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum" or r[ix["Metric Value"]] == "":
            continue
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        rows.append((int(r[ix["ID"]]), short(r[ix["Kernel Name"]]), r[ix["Grid Size"]], r[ix["Block Size"]], ns / 1e3))
    with open(out_list, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "duration_us"])
        for r in rows:
            w.writerow([r[0], r[1], r[2], r[3], f"{r[4]:.3f}"])
    resets = [i for i, r in enumerate(rows) if "reset_kernel" in r[1]]
    bounds = resets + [len(rows)]
    a, b = max(zip(bounds[:-1], bounds[1:]), key=lambda ab: ab[1] - ab[0])   # the longest reset-to-reset segment
    step = rows[a:b]
    tot = sum(r[4] for r in step)
    agg = collections.OrderedDict()
    for r in step:
        n, t = agg.get(r[1], (0, 0.0))
        agg[r[1]] = (n + 1, t + r[4])
    with open(out_shares, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{t:.1f}", f"{t / n:.1f}", f"{100 * t / tot:.2f}"])
    print(f"{len(rows)} launches listed; one step = {len(step)} launches, {tot / 1e3:.1f} ms serialised")


if __name__ == "__main__":
    main(*sys.argv[1:4])
