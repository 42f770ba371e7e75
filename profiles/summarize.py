#!/usr/bin/env python
"""Turn the raw ncu outputs in gpurun_out/ (written by profiles/capture.sh) into the small tracked
summaries under profiles/:
    launches_<tag>.md          per-kernel launch count, device time, share of the step, DRAM bytes
    ncu_<tag>.md               key full-set metrics of the top kernels
    roofline_traffic.json      per-launch DRAM traffic of the dominant kernel (read by bench.py)
usage: python profiles/summarize.py <tag>          (needs `ncu` on PATH to read the .ncu-rep files)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("msntt::", "").replace("ms::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def launches():
    path = os.path.join(OUT, f"launches_{tag}.csv")
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        key = (r["ID"], r["Kernel Name"])
        per.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    for (idx, name), m in per.items():
        t, tu = m.get("gpu__time_duration.sum", (0, "ns"))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1e-6)

        def bytes_of(k):
            v, u = m.get(k, (0, "byte"))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        rows.append((int(idx), short(name), t * scale, bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")))
    return rows


def main():
    rows = launches()
    # the bench runs warm-up step(s) then the timed step: keep the last occurrence block (second half)
    half = len(rows) // 2
    step = rows[half:]
    agg = collections.OrderedDict()
    for _, name, ms, rd, wr in step:
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += rd; a[3] += wr
    total = sum(a[1] for a in agg.values())
    md = [f"# Launch list `{tag}` — one full config-3 step (2^24 x 32 Fp trace, LDE x8, Merkle, constraint eval)",
          "", "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over",
          "`python bench.py --steps 1 --warmup 1 --no-cpu`; the second (timed) step is tabulated.  Times under ncu are",
          "cold-cache and serialised: the SHARES are what is comparable with bench.py's CUDA-event phase times.", "",
          "| kernel | launches | device ms | share | DRAM read GB | DRAM write GB | DRAM GB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
    for name, (cnt, ms, rd, wr) in agg.items():
        md.append(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.1f}% | {rd / 1e9:.2f} | {wr / 1e9:.2f} | {(rd + wr) / 1e6 / max(ms, 1e-9):.0f} |")
    md.append(f"| **total** | {sum(a[0] for a in agg.values())} | {total:.3f} | 100% | | | |")
    # per-launch list of the NTT passes of the LDE (the three longest ntt_pass launches)
    ntt = [r for r in step if r[1].startswith("ntt_pass_kernel") or "ntt_tma_kernel" in r[1]]
    lde = sorted(ntt, key=lambda r: -r[2])[:3]
    md += ["", "## LDE passes (3 launches, all 32 columns x 8 cosets each)", "",
           "| launch | device ms | DRAM read GB | DRAM write GB | algorithmic GB (SURVEY §8d share) |", "|---|---:|---:|---:|---:|"]
    alg = (8 * (1 << 24) + 8 * (1 << 27)) * 32 / 3 / 1e9
    for r in sorted(lde):
        md.append(f"| {r[1]} #{r[0]} | {r[2]:.3f} | {r[3] / 1e9:.2f} | {r[4] / 1e9:.2f} | {alg:.2f} |")
    open(os.path.join(ROOT, "profiles", f"launches_{tag}.md"), "w").write("\n".join(md) + "\n")
    traffic = {"tag": tag, "lde_dram_bytes_per_launch": sum(r[3] + r[4] for r in lde) / max(len(lde), 1),
               "lde_algorithmic_bytes_per_launch": alg * 1e9,
               "source": f"profiles/launches_{tag}.md (ncu dram__bytes_read.sum + dram__bytes_write.sum, full-size step)"}
    old = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(old):           # keep hand-entered annotations (ncu pipe utilisation of the full-set capture)
        try:
            prev = json.load(open(old))
            for k in prev:
                traffic.setdefault(k, prev[k])
        except Exception:
            pass
    json.dump(traffic, open(old, "w"), indent=1)

    # full-set captures
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "sm__inst_executed_pipe_tma.sum",
            "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    md = [f"# ncu --set full summaries `{tag}`" + (" (NTT kernels: 2^24-point transforms, 4 columns x 8 cosets, profiles/exp_ntt_r02.py prof; hash / "
          "evaluator kernels: 2^22-row instance of the bench step)" if tag.startswith("r02") else " (2^22-row instance of the same pipeline)"), ""]
    for rep in (f"prof_ntt_lde_{tag}.ncu-rep", f"prof_hash_eval_{tag}.ncu-rep"):
        path = os.path.join(OUT, rep)
        if not os.path.exists(path):
            continue
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        md += [f"## {rep}", ""]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            md.append(f"### `{short(d['Kernel Name'])}`  (ID {d['ID']})")
            md.append("")
            md.append("| metric | value | unit |")
            md.append("|---|---:|---|")
            for w in want:
                if w in d:
                    md.append(f"| {w} | {d[w]} | {units[hdr.index(w)]} |")
            md.append("")
    open(os.path.join(ROOT, "profiles", f"ncu_{tag}.md"), "w").write("\n".join(md) + "\n")
    print("wrote profiles/launches_%s.md, profiles/ncu_%s.md, profiles/roofline_traffic.json" % (tag, tag))


if __name__ == "__main__":
    main()
