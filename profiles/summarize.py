#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py kernel gpurun_out/prof_x.ncu-rep profiles/r1_link_pcg2.md [pairs]
  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * mult.get(unit, 1)


def kernel(rep, out_md, pairs=None):
    rows = ncu_csv(rep, "raw")
    hdr, unit, val = rows[0], rows[1], rows[2]
    m = {h: (val[i], unit[i]) for i, h in enumerate(hdr)}
    name = m.get("Kernel Name", ("?", ""))[0]
    lines = [f"# ncu --set full: `{name}`", "", f"source report: `{os.path.basename(rep)}` (not tracked)", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(m[h][0])))
            except ValueError:
                pass
    stalls.sort(key=lambda t: -t[1])
    lines += ["", "warps stalled per issue-active cycle, by reason (smsp__average_warps_issue_stalled_*):", ""]
    for h, v in stalls[:8]:
        lines.append(f"* {h}: {v:.2f}")
    for k in ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"):
        if k in m:
            lines.append(f"* {k}: {m[k][0]} %")
    dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    dur_ms = float(m["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[m["gpu__time_duration.sum"][1]]
    inst = float(m["smsp__inst_executed.sum"][0])
    lines += ["", f"DRAM traffic per launch: {dram / 1e9:.3f} GB; duration under ncu (cold, serialised): {dur_ms:.2f} ms"]
    if pairs:
        pairs = float(pairs)
        lines += [f"pairs per launch: {pairs:.4g} -> {inst * 32 / pairs:.1f} thread-instructions per (record, candidate) pair, "
                  f"{inst / (pairs / 32):.1f} warp-instructions per 32-candidate step"]
    # hot loop opcode mix from the source page
    src = ncu_csv(rep, "source")
    sh = src[1]
    ci = {h: i for i, h in enumerate(sh)}
    body = [r for r in src[2:] if len(r) > 10]
    if body and "Instructions Executed" in ci:
        counts = sorted(int(r[ci["Instructions Executed"]]) for r in body)
        # the main loop = the most populated band of equal execution counts (the mbarrier spin loop has higher
        # counts but only a handful of instructions)
        import statistics
        big = [c for c in counts if c > 0.02 * counts[-1]]
        mode = statistics.median(big) if big else counts[-1]
        hot = [r for r in body if 0.7 * mode <= int(r[ci["Instructions Executed"]]) <= 1.4 * mode]
        op = collections.Counter()
        for r in hot:
            t = r[ci["Source"]].split()
            o = (t[1] if t[0].startswith("@") else t[0]).rstrip(";").split(".")[0]
            op[o] += 1
        lines += ["", f"main loop ({len(hot)} SASS instructions with the modal execution count), opcode mix:", "",
                  ", ".join(f"{k} {v}" for k, v in op.most_common(16))]
        sass = " ".join(r[ci["Source"]] for r in body)
        lines += ["", "Blackwell/Hopper async-copy evidence in SASS: " +
                  ", ".join(k for k in ("UBLKCP", "SYNCS", "UTMALDG") if k in sass)]
    open(out_md, "w").write("\n".join(lines) + "\n")
    tj = os.path.join(os.path.dirname(out_md), "traffic.json")
    traffic = json.load(open(tj)) if os.path.exists(tj) else {}
    traffic[os.path.basename(out_md).replace(".md", "")] = dram
    json.dump(traffic, open(tj, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


def launches(csv_path, out_md):
    rows = [r for r in csv.reader(open(csv_path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[ci["Metric Unit"]]
        v = float(r[ci["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        per[r[ci["Kernel Name"]].split("(")[0][:70]].append(v)
    tot = sum(sum(v) for v in per.values())
    lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)",
             "", f"source: `{os.path.basename(csv_path)}`; {sum(len(v) for v in per.values())} launches, {tot:.1f} ms total", "",
             "| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v):.3f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.4f} |")
    open(out_md, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "kernel":
        kernel(*sys.argv[2:])
    else:
        launches(*sys.argv[2:])
