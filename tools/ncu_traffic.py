"""Builds profiles/traffic.json -- DRAM bytes per sample of the MLP kernels, from the committed `ncu --set full` raw pages
(profiles/*_raw.csv: metric,unit,value as written by `ncu -i X.ncu-rep --page raw --csv` after transposition).  bench.py
scales these to the average launch of its own run for `roofline.traffic` (it cannot read DRAM counters itself); re-run
after every new capture:

    python tools/ncu_traffic.py mlp_forward=profiles/r2a_tc_raw.csv:524288 mlp_forward_stash=...:262144 ...

A part may name the kernels it counts, `path@regex:samples` (all launches of the page whose name matches are summed;
`samples` = the samples those launches processed together), and parts joined by `+` add their per-sample figures:

    python tools/ncu_traffic.py render_forward=profiles/r2i_fwd_raw.csv@mlp_fwd_tc:40960000 \
        mlp_backward=profiles/r2i_train_raw.csv@mlp_bwd_tc:786432+profiles/r2i_train_raw.csv@dw_tc:524288
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path, kernel=None):
    """Both shapes of a raw page: transposed (metric,unit,value per line) or ncu's own (names / units / one row per launch;
    `kernel` = regex selecting the launches to sum, default: the first launch only, as the earlier captures were used)."""
    import re
    rows = list(csv.reader(open(path)))
    tot = 0.0
    if rows and rows[0] and rows[0][0] == "ID":
        names, units = rows[0], rows[1]
        kn = names.index("Kernel Name")
        picked = rows[2:3] if kernel is None else [r for r in rows[2:] if re.search(kernel, r[kn])]
        for vals in picked:
            for n, u, v in zip(names, units, vals):
                if n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(v.replace(",", "")) * UNITS[u]
        return tot
    for row in rows:
        if len(row) >= 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(row[2]) * UNITS[row[1]]
    return tot


def main():
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(out_path)) if os.path.isfile(out_path) else {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=")
        if "@" in rest:
            per_sample, n = 0.0, 0
            for part in rest.split("+"):
                spec, samples = part.rsplit(":", 1)
                path, kernel = spec.split("@")
                per_sample += dram_bytes(os.path.join(ROOT, path), kernel) / int(samples)
                n = max(n, int(samples))
            out[key] = {"dram_bytes_per_sample": per_sample, "samples_per_captured_launch": n, "source": rest}
            continue
        paths, samples = rest.rsplit(":", 1)
        b = sum(dram_bytes(os.path.join(ROOT, p)) for p in paths.split("+"))
        out[key] = {"dram_bytes_per_sample": b / int(samples), "samples_per_captured_launch": int(samples),
                    "source": paths}
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
