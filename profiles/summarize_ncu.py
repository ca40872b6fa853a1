#!/usr/bin/env python
"""Turn gpurun_out/<tag>_{launches.csv,full.ncu-rep,plain.log} into committed summaries profiles/<tag>.md.

    python profiles/summarize_ncu.py <tag> [<tag> ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]


def raw_page(tag, rep):
    """The raw metric page of the capture: the CSV exported on the GPU box (gpu_profile.sh) when it is there and not older
    than a report of the same tag, else exported here from the report."""
    raw = os.path.join(OUT, f"{tag}_raw.csv")
    if os.path.exists(raw) and (not os.path.exists(rep) or os.path.getmtime(raw) >= os.path.getmtime(rep)):
        return open(raw).read()
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def launches(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows:
        name = r[ik].split("(")[0][:90]
        t = float(r[iv].replace(",", ""))
        n, s = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, s + t)
    total = sum(s for _, s in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {n} | {s / 1e6:.3f} | {100 * s / total:.1f} % |")
    return "\n".join(out)


def full(tag):
    rep = os.path.join(OUT, f"{tag}_full.ncu-rep")
    txt = raw_page(tag, rep)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    out = [f"kernel: `{d['Kernel Name'][0]}`", "", "| metric | value | unit |", "|---|---:|---|"]
    for k in KEYS:
        if k in d:
            out.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
    return "\n".join(out)


def traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum of the profiled launch, in bytes (None if absent)."""
    rep = os.path.join(OUT, f"{tag}_full.ncu-rep")
    txt = raw_page(tag, rep)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k not in hdr:
            return None
        i = hdr.index(k)
        tot += float(vals[i].replace(",", "")) * mult.get(units[i], 1)
    sys.path.insert(0, ROOT)
    from q_learning_with_hjb_b200.build import source_hash
    kernel = vals[hdr.index("Kernel Name")][:80]
    return {"dram_bytes_per_launch": tot, "kernel": kernel,
            "kernel_source_hash": source_hash("rollout" if "rollout" in kernel else "vhjb"),
            "source": f"profiles/{tag}.md (ncu --set full, one launch)"}


def main():
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for tag in sys.argv[1:]:
        t = traffic(tag)
        if t:
            table[tag] = t
    with open(tpath, "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    for tag in sys.argv[1:]:
        plain = open(os.path.join(OUT, f"{tag}_plain.log")).read().strip().splitlines()[-1]
        try:
            line = json.loads(plain)
            head = json.dumps({k: line[k] for k in ("metric", "value", "unit", "ms_per_step", "config") if k in line})
        except ValueError:
            head = plain[:400]
        md = [f"# {tag}", "", "Command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vhjb ...` "
              "(see profiles/gpu_profile.sh); numbers printed under ncu are NOT bench values.", "",
              "Plain run (no profiler):", "", "```", head, "```", "",
              "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: shares, not absolutes)",
              "", launches(tag), "", "## Top kernel (`ncu --set full --clock-control none --import-source on`, one launch)", "",
              full(tag), ""]
        with open(os.path.join(ROOT, "profiles", f"{tag}.md"), "w") as fh:
            fh.write("\n".join(md))
        print("wrote", f"profiles/{tag}.md")


if __name__ == "__main__":
    main()
