"""Rebuild profiles/traffic.json (the DRAM bytes bench.py reports next to the roofline) from the ncu summaries committed in
this directory: dram__bytes_read.sum + dram__bytes_write.sum per launch, first launch of each kernel in each file."""
import json, os, re

HERE = os.path.dirname(os.path.abspath(__file__))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def kernels(fname):
    """[(kernel name, read + write bytes)] in launch order"""
    out, name, rd = [], None, None
    for line in open(os.path.join(HERE, fname)):
        if line.startswith("kernel:"):
            name, rd = line[len("kernel:"):].strip(), None
        m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m:
            v = float(m.group(2)) * UNIT[m.group(3)]
            if m.group(1) == "read":
                rd = v
            else:
                out.append((name, int(round(rd + v))))
    return out


def first(ks, pat):
    return next(b for n, b in ks if re.search(pat, n))


def total(ks, pat):
    return sum(b for n, b in ks if re.search(pat, n))


c2 = kernels("r2_c2_exact_full.txt")
stage = {
    "coherence_probe": first(c2, "coherence_probe"),
    "tile_hist": first(c2, "tile_hist"),
    "tile_offsets": first(c2, "bin_totals") + first(c2, "tile_scan") + first(c2, "bin_cursors"),
    "tile_scatter": first(c2, "tile_scatter"),
    "eval_brick": first(c2, "eval_brick"),
    "eval_deferred": first(c2, "eval_deferred"),
    "unpack (unpermute_kernel)": first(c2, "unpermute"),
    "eval_kernel (idle: batch not coherent)": first(c2, r"eval_kernel"),
}
fast_brick = first(kernels("r2_c2_fast_full.txt"), "eval_brick")
fast_stage = dict(stage, eval_brick=fast_brick)
c5 = kernels("r2_c5_full.txt")
pf3e, pf3f, pf2 = kernels("r2_c3_prefilter_exact_full.txt"), kernels("r2_c3_prefilter_fast_full.txt"), kernels("r2_c2_prefilter_full.txt")
doc = {
    "source": "ncu --set full --clock-control none (profiles/ncu_round2.sh), dram__bytes_read.sum + dram__bytes_write.sum per launch; "
              "summaries in profiles/r2_*_full.txt; this file is rebuilt from them by profiles/make_traffic.py",
    "eval_brick_exact_bytes_per_launch": stage["eval_brick"],
    "eval_brick_fast_bytes_per_launch": fast_brick,
    "c2_step_exact_by_stage": stage,
    "c2_step_fast_by_stage": fast_stage,
    "c2_step_exact_bytes": sum(stage.values()),
    "c2_step_fast_bytes": sum(fast_stage.values()),
    "c1_plain_bytes_per_1e8": first(kernels("r2_c1_plain_full.txt"), "eval_kernel"),
    "c3_plain_blocked_bytes_per_1e8": first(kernels("r2_c3_plain_blocked_full.txt"), "eval_kernel"),
    "c4_plain_blocked_periodic_bytes_per_1e8": first(kernels("r2_c4_plain_blocked_full.txt"), "eval_kernel"),
    "c5_restrict_level0_bytes": first(c5, "restrict3_march"),
    "c5_prolong_level0_bytes": first(c5, "prolong3_march"),
    "c3_prefilter_exact_bytes": total(pf3e, "line_solve"),
    "c3_prefilter_fast_tma_sweep_bytes": first(pf3f, "fast_tma"),
    "c2_prefilter_bytes": total(pf2, "pad1|line_solve"),
    "sources": {
        "c2 step / eval_brick exact": "profiles/r2_c2_exact_full.txt", "eval_brick fast": "profiles/r2_c2_fast_full.txt",
        "c1": "profiles/r2_c1_plain_full.txt", "c3": "profiles/r2_c3_plain_blocked_full.txt", "c4": "profiles/r2_c4_plain_blocked_full.txt",
        "c5": "profiles/r2_c5_full.txt", "c3 prefilter": "profiles/r2_c3_prefilter_exact_full.txt, profiles/r2_c3_prefilter_fast_full.txt",
        "c2 prefilter": "profiles/r2_c2_prefilter_full.txt",
    },
}
json.dump(doc, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
print(json.dumps(doc, indent=1))
