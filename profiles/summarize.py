#!/usr/bin/env python
"""Turn the ncu artefacts that profiles/capture.sh leaves in gpurun_out/ into the tracked summaries:

    python profiles/summarize.py r1      ->  profiles/launches_r1.md, profiles/counters_r1.md

launches: per-kernel launch count, total device time and SHARE of the profiled run (per-launch
times are cold-cache and serialised; only shares are comparable with bench.py).
counters: one row per profiled launch of our kernels from the `--set full` capture.
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out")
out = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("kgat::<unnamed>::", "kgat::").replace("at::native::", "at::")
    return name[:90]


rows = list(csv.reader(open(os.path.join(src, "launches_%s.csv" % tag))))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(r[ui], v)
    a = agg.setdefault(short(r[ki]), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith("kgat::"))
with open(os.path.join(out, "launches_%s.md" % tag), "w") as f:
    f.write("# ncu launch list, round %s\n\n" % tag)
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py "
            "--steps 2 --warmup 1 --no-cpu-baseline` (profiles/capture.sh).\n")
    f.write("Covers graph build, the eager + graph-replayed + e2e steps and the evaluation. %d launches, %.1f us total, "
            "%.1f%% of the time in kgat:: kernels.\n\n" % (len(data), tot, 100 * ours / tot))
    f.write("| kernel | launches | total us | us / launch | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))

COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"),
        ("dram__bytes_write.sum", "dram wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed.avg.per_cycle_active", "IPC"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_registers", "occ lim regs"), ("launch__occupancy_limit_shared_mem", "occ lim smem")]


def bench_key(name):
    """ncu kernel name -> the key bench.py uses for the same launch."""
    m = re.search(r"layer_(fwd|bwd)_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)[,>]", name)
    if m:
        return "layer_%s_%sx%s" % m.groups()
    m = re.search(r"(input_bwd|csc_gather)_kernel<\(?(?:int\))?(\d+)[,>]", name)
    if m:
        return "%s_%s" % m.groups()
    for pat, key in (("att_logits", "attention_logits"), ("segment_softmax", "segment_softmax"),
                     ("gather_weights", "edge_weights_to_csc"), ("score_topk_f16", "score_topk_f16"),
                     ("topk_merge", "topk_merge"), ("eval_split", "eval_split"), ("bpr_fwd", "bpr_loss_fwd"),
                     ("bpr_bwd", "bpr_loss_bwd"), ("transr_main", "transr_main"), ("transr_plan", "transr_plan")):
        if pat in name:
            return key
    return None


def load_rep(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


reps = [(sfx, os.path.join(src, "prof_%s%s.ncu-rep" % (tag, sfx))) for sfx in ("", "_eval", "_kg")]
reps = [(sfx, r) for sfx, r in reps if os.path.exists(r)]
if reps:
    import json
    byte_units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    time_units = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    kernels = collections.OrderedDict()
    with open(os.path.join(out, "counters_%s.md" % tag), "w") as f:
        f.write("# ncu --set full counters, round %s\n\nOne row per profiled launch (`ncu --set full --clock-control "
                "none --import-source on`, profiles/capture.sh).  Sections: the bench step (amazon-book shape), the "
                "evaluation kernels, one TransR iteration.\n" % tag)
        for sfx, rep in reps:
            hdr, units, data = load_rep(rep)
            idx = [(hdr.index(c) if c in hdr else None, t) for c, t in COLS]
            f.write("\n## prof_%s%s.ncu-rep\n\n" % (tag, sfx))
            f.write("| " + " | ".join(t for _, t in idx) + " |\n|" + "---|" * len(idx) + "\n")
            for r in data:
                cells = []
                for i, t in idx:
                    if i is None:
                        cells.append("n/a")
                    elif t == "kernel":
                        cells.append("`%s`" % short(r[i]))
                    else:
                        cells.append((r[i] + " " + units[i]).strip())
                f.write("| " + " | ".join(cells) + " |\n")

            def num(r, col, scale_units):
                i = hdr.index(col)
                return float(r[i].replace(",", "")) * scale_units.get(units[i], 1.0)

            ki = hdr.index("Kernel Name")
            for r in data:
                k = bench_key(r[ki])
                if k and k not in kernels:
                    kernels[k] = {"dram_bytes": num(r, "dram__bytes_read.sum", byte_units)
                                  + num(r, "dram__bytes_write.sum", byte_units),
                                  "ncu_time_us": num(r, "gpu__time_duration.sum", time_units)}
    # measured DRAM traffic per launch (first profiled launch of each kernel) -> what bench.py reads for
    # roofline.traffic
    json.dump({"source": "ncu --set full --clock-control none, profiles/capture.sh %s (amazon-book shape)" % tag,
               "kernels": kernels}, open(os.path.join(out, "traffic_%s.json" % tag), "w"), indent=1)
print("wrote", sorted(p for p in os.listdir(out) if tag in p))
