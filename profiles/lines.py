#!/usr/bin/env python
"""Per-source-line cost of one kernel: joins the SASS page of an ncu report (instructions executed, stall samples)
with the line table of the object file (nvdisasm -g), by instruction order.

  python profiles/lines.py gpurun_out/prof_x.ncu-rep dblink_b200/build/dbl_engine.o k_link_pruned [top]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
        cub = glob.glob(os.path.join(d, "*.cubin"))[0]
        out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    lines, cur, inside = [], None, False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    return lines


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ci = {k: i for i, k in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) > 20]
    lt = sass_lines(obj, kernel)
    if len(lt) != len(body):
        print(f"warning: {len(lt)} instructions in the object, {len(body)} in the report", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for r, where in zip(body, lt):
        a = agg[where]
        a[0] += int(r[ci["Instructions Executed"]])
        a[1] += int(r[ci["# Samples"]])
        for sc in stall_cols:
            a[2][sc] += int(r[ci[sc]] or 0)
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    print(f"{kernel}: {ti} warp-instructions, {ts} samples")
    print("| file:line | instr % | samples % | top stalls |")
    print("|---|---|---|---|")
    for where, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = ", ".join(f"{k[6:]} {v}" for k, v in a[2].most_common(2))
        print(f"| {where[0]}:{where[1]} | {100 * a[0] / ti:.1f} | {100 * a[1] / ts:.1f} | {st} |" if where else f"| ? | {100*a[0]/ti:.1f} | {100*a[1]/ts:.1f} | {st} |")


if __name__ == "__main__":
    main()
