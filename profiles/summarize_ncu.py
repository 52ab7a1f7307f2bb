"""Turns ncu output brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python profiles/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt
    python profiles/summarize_ncu.py traffic  gpurun_out/traffic.ncu-rep "<workload name>" [batches per step] > profiles/trace_ncu_summary.json
        (a capture of every k_primary / k_trace / k_tail launch of `bench.py --steps 1 --warmup 1`: DRAM bytes
         of the traversal launches of the LAST step, which bench.py reports as roofline.traffic)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("dsrt::<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        order.append((name, v))
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(order)} launches, {tot:.1f} us summed (ncu: cold-cache, serialised -- compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} n={v[0]:5d} total={v[1]:11.1f} us share={v[1] / tot:6.3f}")
    # shares within the render step only (this library's frame kernels), to set beside bench.py's
    # kernel_ms_per_step / ms_per_step
    step = ("k_batch_begin", "k_primary", "k_shade", "k_advance", "k_trace", "k_tail", "k_resolve", "k_tonemap")
    sagg = {k: v for k, v in agg.items() if k.split("<")[0] in step}
    stot = sum(v[1] for v in sagg.values())
    if stot > 0:
        print(f"# render-step kernels only: {stot:.1f} us summed")
        for k, v in sorted(sagg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {k:46s} n={v[0]:5d} total={v[1]:11.1f} us share_of_step={v[1] / stot:6.3f}")
        trav = sum(v[1] for k, v in sagg.items() if k.split("<")[0] in ("k_primary", "k_trace", "k_tail"))
        print(f"  traversal kernels (k_primary + k_trace + k_tail) share_of_step={trav / stot:6.3f}")
    print("# first 60 launches in order")
    for n, v in order[:60]:
        print(f"  {n:46s} {v:10.1f} us")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")][:60] for r in data]
    print("# launches:", " | ".join(names))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:88s} {units[i]:16s} " + " | ".join(r[i] for r in data))


def traffic(path, workload, batches=1):
    """batches = k_primary launches per step (a step whose samples do not fit one batch has several)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                     "l1tex__m_xbar2l1tex_read_bytes.sum")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1.0)
    steps, cur = [], None
    for r in data:
        name = r[col["Kernel Name"]]
        short = "k_primary" if "k_primary" in name else "k_trace" if "k_trace" in name else "k_tail" if "k_tail" in name else None
        if short is None:
            continue
        if short == "k_primary":
            cur = []
            steps.append(cur)
        if cur is None:
            continue
        cur.append({"kernel": short, "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                    "l2_to_sm_read_bytes": val(r, "l1tex__m_xbar2l1tex_read_bytes.sum"), "ms": val(r, "gpu__time_duration.sum")})
    last = [x for st in steps[-batches:] for x in st]
    res = {"workload": workload, "source": path.split("/")[-1],
           "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, default cache control (every launch starts with "
                   "cold caches): an upper bound on the DRAM traffic of the same launches inside an un-profiled step",
           "launches_in_step": len(last),
           "dram_bytes_per_step": sum(x["dram_read"] + x["dram_write"] for x in last),
           "dram_read_bytes_per_step": sum(x["dram_read"] for x in last),
           "dram_write_bytes_per_step": sum(x["dram_write"] for x in last),
           "l2_to_sm_read_bytes_per_step": sum(x["l2_to_sm_read_bytes"] for x in last),
           "ncu_ms_per_step": sum(x["ms"] for x in last),
           "per_launch": last}
    print(json.dumps(res, indent=1))


def lines(path, kernel_regex, skip="0", top="45"):
    """Per source line of one launch: share of the executed warp instructions, average active threads per
    instruction (lane utilisation) and share of the stall samples -- from `ncu --import-source on` reports of code
    built with -lineinfo.    python profiles/summarize_ncu.py lines <rep> <kernel regex> [launch-skip] [top n]"""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          f"regex:{kernel_regex}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr, acc, fn = None, None, {}, ""
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r and r[0] == "Function Name":
            fn = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10 or r[2] != "-":       # keep the per-line aggregates, skip the per-SASS rows
            continue
        try:
            ie, te, smp = float(r[7]), float(r[8]), float(r[6])
        except ValueError:
            continue
        o = acc.setdefault((cur, int(r[0])), [0.0, 0.0, 0.0, r[1].strip()[:110]])
        o[0] += ie; o[1] += te; o[2] += smp
    tot = sum(o[0] for o in acc.values()) or 1.0
    tott = sum(o[1] for o in acc.values())
    tots = sum(o[2] for o in acc.values()) or 1.0
    print(f"# {fn}: {tot:.4g} warp instructions, {tott / tot:.2f} threads per instruction, {int(tots)} stall samples")
    print("# file line | % of instructions | threads per instruction | % of samples | source")
    for (f, l), o in sorted(acc.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        print(f"{f:14s} {l:5d} {100 * o[0] / tot:6.2f}% {o[1] / max(o[0], 1):5.1f} {100 * o[2] / tots:6.2f}%  {o[3]}")


if __name__ == "__main__":
    if sys.argv[1] == "lines":
        lines(*sys.argv[2:6])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
