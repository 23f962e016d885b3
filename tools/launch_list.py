"""Aggregate an ncu launch list (gpu__time_duration.sum, --csv) of `bench.py --steps 1 --warmup 3 --no-graph ...` into shares per kernel for
ONE step: the bench launches the same kernels every step, so the list is cut into equal chunks and the 4th chunk (the timed step) is reported.

    python tools/launch_list.py gpurun_out/launches_r02h_nvfp4.csv [steps_in_list]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = rows[hdr]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    ui = cols.index("Metric Unit")
    launches = []
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1e-3)
        launches.append((r[ki], v))
    # find the period: the smallest n such that the kernel-name sequence repeats
    names = [re.sub(r"\(.*", "", n) for n, _ in launches]
    period = None
    for n in range(200, len(names) // 2 + 1):
        if names[:n] == names[n:2 * n]:
            period = n
            break
    if period is None:
        period = len(names) // (int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    step = launches[3 * period:4 * period] if len(launches) >= 4 * period else launches[-period:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        k = re.sub(r"^void |<unnamed>::|nb200::|\(anonymous namespace\)::", "", re.sub(r"\(.*", "", n))
        agg[k][0] += 1
        agg[k][1] += v
    total = sum(v for _, v in agg.values())
    print(f"launches in the list {len(launches)}, per step {period}, step total under ncu {total / 1e3:.2f} ms\n")
    print("| share | launches | avg µs | kernel |\n|---|---|---|---|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * v / total:.1f} % | {c} | {v / c:.1f} | `{k}` |")


if __name__ == "__main__":
    main()
