#!/usr/bin/env python
"""Turn the `.ncu-rep` files of tools/profile_round.sh (gpurun_out/<round>_prof_*.ncu-rep) into the
committed evidence: profiles/<round>_prof_<name>.raw.csv (ncu's raw page), profiles/<round>_ncu_summary.md
and profiles/ncu_summary.json (the per-launch DRAM traffic bench.py quotes in `roofline.traffic`).

    python tools/ncu_summarize.py r1
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "dram__bytes.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]
# capture name -> (title, key in ncu_summary.json or None)
CAPTURES = {
    "agg": ("aggregation kernel, reddit-shaped R-MAT, F=128 (headline workload)", "reddit_gcn_layer_128"),
    "agg_uniform": ("same kernel, uniformly random sources (cache-hostile context run)", "reddit_gcn_layer_128_uniform"),
    "fixup": ("carry fix-up kernels of the headline workload", None),
    "dense": ("warp-specialised tcgen05 3xTF32 combination (W resident), 232,965 x 128 x 128", None),
    "agg_products": ("aggregation kernel, products-shaped R-MAT, F=256", "products_gcn_layer_256"),
    "dense_stream": ("warp-specialised tcgen05 3xTF32 combination (W streamed), 2,449,029 x 256 x 256", None),
    "agg_proteins": ("aggregation kernel, proteins-shaped R-MAT, F=64", "proteins_gcn_layer_64"),
    "agg_arxiv": ("aggregation kernel, arxiv-shaped R-MAT, F=32", "arxiv_gcn_layer_32"),
    "gat": ("fused GAT aggregation, proteins-shaped R-MAT, F=64", None),
    "agg_rmat26": ("aggregation kernel, R-MAT scale 26 (67 M vertices / 1.07 G edges), F=64, one GPU", "rmat26_gcn_agg_64"),
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def raw_page(rep):
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout


def parse(raw):
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, out = rows[hdr], rows[hdr + 1], []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {}
        for nm, u, v in zip(names, units, r):  # plain columns win over the "<unit>.Triage*." duplicates
            parts = nm.split(".", 2)
            triage = len(parts) == 3 and parts[1].startswith("Triage")
            if v == "" and nm != "Kernel Name":
                continue
            if triage:
                d.setdefault(parts[2], (v, u))
            else:
                d[nm] = (v, u)
        out.append(d)
    return out


def num(vu):
    v, u = vu
    return float(v.replace(",", "")) * UNIT_SCALE.get(u, 1)


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    md = ["# Round %s ncu summaries (B200, `--clock-control none`)\n" % rnd[1:],
          "Source reports: `gpurun_out/%s_prof_*.ncu-rep` (scratch), produced by `tools/profile_round.sh %s` "
          "(`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 3 -c 1 python bench.py --steps 2 "
          "--warmup 3 [--workload ...]`); raw pages committed as `profiles/%s_prof_*.raw.csv`; this file written by "
          "`tools/ncu_summarize.py`.\n" % (rnd, rnd, rnd)]
    jpath = os.path.join(dst, "ncu_summary.json")
    summ = json.load(open(jpath)) if os.path.exists(jpath) else {}
    for name, (title, key) in CAPTURES.items():
        rep = os.path.join(src, "%s_prof_%s.ncu-rep" % (rnd, name))
        rawp = os.path.join(dst, "%s_prof_%s.raw.csv" % (rnd, name))
        box_raw = rep.replace(".ncu-rep", ".raw.csv")
        if os.path.exists(rep):
            raw = raw_page(rep)
            open(rawp, "w").write(raw)
        elif os.path.exists(box_raw) and os.path.getsize(box_raw) > 0:
            raw = open(box_raw).read()
            open(rawp, "w").write(raw)
        elif os.path.exists(rawp):
            raw = open(rawp).read()
        else:
            continue
        for k in parse(raw):
            kn = k["Kernel Name"][0]
            md.append("## %s — `%s`\n" % (title, kn))
            md.append("| metric | value |\n|---|---|")
            for m in METRICS:
                if m in k:
                    md.append("| %s | %s %s |" % (m, k[m][0], k[m][1]))
            md.append("")
            if key:
                rd, wr = num(k["dram__bytes_read.sum"]), num(k["dram__bytes_write.sum"])
                summ[key] = {"kernel": kn, "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                             "duration_ms": num(k["gpu__time_duration.sum"]),
                             "l1_hit_pct": num(k["l1tex__t_sector_hit_rate.pct"]), "l2_hit_pct": num(k["lts__t_sector_hit_rate.pct"]),
                             "l1tex_throughput_pct": num(k["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]),
                             "lts_throughput_pct": num(k["lts__throughput.avg.pct_of_peak_sustained_elapsed"]),
                             "dram_throughput_pct": (num(k["dram__throughput.avg.pct_of_peak_sustained_elapsed"])
                                                     if "dram__throughput.avg.pct_of_peak_sustained_elapsed" in k else None),
                             "dram_GBps": round((rd + wr) / num(k["gpu__time_duration.sum"]) / 1e6, 1),
                             "instructions": num(k["smsp__inst_executed.sum"]),
                             "source": "profiles/%s_prof_%s.raw.csv" % (rnd, name)}
    open(os.path.join(dst, "%s_ncu_summary.md" % rnd), "w").write("\n".join(md) + "\n")
    json.dump(summ, open(jpath, "w"), indent=1)
    print("wrote", "%s_ncu_summary.md" % rnd, "and ncu_summary.json with keys", sorted(summ))


if __name__ == "__main__":
    main()
