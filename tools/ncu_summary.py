"""Summarise an ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --clock-control none --csv --log-file X ...`) into a per-kernel table (markdown) and
a per-family DRAM-traffic JSON (what bench.py reports as roofline.traffic).

  python tools/ncu_summary.py launches.csv[.gz] out_summary.md out_traffic.json "<command that was profiled>" [passes]

`passes` (optional): keep only the launches up to and including the merge of the passes-th `k_classify` launch,
i.e. exactly one benchmark step (C2: 101 iterations + 1 nested pass = 102) when the capture ran on into the next.
"""
import csv
import gzip
import io
import json
import re
import sys
from collections import defaultdict

FAMILIES = [("merge", re.compile(r"k_merge_")), ("sign", re.compile(r"k_sign")),
            ("group", re.compile(r"k_radix|k_heads|k_classify|k_size_hist|k_scan_single")),
            ("compact", re.compile(r"k_alive"))]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:60]


def main():
    src, out_md, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    passes = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    raw = gzip.open(src, "rt").read() if src.endswith(".gz") else open(src).read()
    lines = [ln for ln in raw.splitlines() if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    per_launch = defaultdict(dict)
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] else 0.0
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_launch[(r["ID"], r["Kernel Name"])][r["Metric Name"]] = v
    if passes:  # cut the capture after one step: launch ids are chronological
        ids = sorted(per_launch, key=lambda k: int(k[0]))
        seen, cut = 0, None
        for k in ids:
            if "k_classify" in k[1]:
                seen += 1
                if seen == passes + 1:
                    cut = int(k[0])
                    break
        if cut is not None:
            # the (passes+1)-th classify belongs to the next step, and so do its sign/sort launches just before it
            last_merge = max(int(k[0]) for k in ids if int(k[0]) < cut and "k_merge" in k[1])
            per_launch = {k: v for k, v in per_launch.items() if int(k[0]) <= last_merge + 2}
    kern = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for (_id, name), m in per_launch.items():
        k = kern[short(name)]
        k[0] += 1
        k[1] += m.get("gpu__time_duration.sum", 0.0)
        k[2] += m.get("dram__bytes_read.sum", 0.0)
        k[3] += m.get("dram__bytes_write.sum", 0.0)
    total_ms = sum(v[1] for v in kern.values())
    lib = {k: v for k, v in kern.items() if k.startswith("k_") or k.startswith("sign_umma::")}
    lib_ms = sum(v[1] for v in lib.values())
    fam = {}
    for fname, rx in FAMILIES:
        sel = [v for k, v in lib.items() if rx.search(k)]
        fam[fname] = {"ms": sum(v[1] for v in sel), "share": sum(v[1] for v in sel) / lib_ms if lib_ms else 0.0,
                      "dram_GB": sum(v[2] + v[3] for v in sel) / 1e9, "launches": sum(v[0] for v in sel)}
    with open(out_md, "w") as f:
        f.write("# ncu launch list\n\nCommand (after the same command had exited 0 without ncu):\n\n```\n%s\n```\n\n" % cmd)
        f.write("Times are per-launch, serialised and cold-cache: only the SHARES are comparable with the live CUDA-event numbers of `bench.py`.\n\n")
        f.write("| kernel | launches | ms | share | DRAM read GB | DRAM write GB |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:24]:
            f.write("| `%s` | %d | %.2f | %.1f %% | %.2f | %.2f |\n" % (k, v[0], v[1], 100 * v[1] / total_ms, v[2] / 1e9, v[3] / 1e9))
        f.write("\nFamilies (library kernels only, i.e. without torch's input synthesis; %d launches, %.1f ms):\n\n" % (sum(v[0] for v in lib.values()), lib_ms))
        for fname, d in fam.items():
            f.write("* %s: %.1f ms = %.1f %% of the library's kernel time, DRAM traffic %.0f GB over %d launches\n" % (
                fname, d["ms"], 100 * d["share"], d["dram_GB"], d["launches"]))
    json.dump({"source": "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, summed per family over the captured step)" % src,
               "families": fam}, open(out_json, "w"), indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    main()
