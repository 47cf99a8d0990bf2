"""Per-kernel launch counts, total time and share from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python profiles/launch_shares.py gpurun_out/launches.csv ["title"]
"""
import collections
import csv
import re
import sys


def main(path, title=None):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"imc::", "", name)
        if "cub::" in name or "DeviceRadixSort" in name:
            name = "cub::DeviceRadixSort*"
        t = float(r[iv].replace(",", ""))
        unit = rows[1][h.index("Metric Unit")] if "Metric Unit" in h else "ns"
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    tot = sum(v[1] for v in agg.values())
    if title:
        print(f"### {title}\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] * scale:.1f} | {100 * v[1] / tot:.1f}% |")
    print(f"\ntotal device time: {tot * scale / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
