#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full --import-source on) into a small text summary for profiles/:
key launch metrics per captured kernel + the source lines holding most warp-stall samples.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep <mangled kernel symbol> <cubin name inside the .so> > profiles/x.txt

The line attribution joins ncu's SASS-level samples with `nvdisasm -g` of the shipped cubin (same instruction order)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "lg_throttle", "mio_throttle", "branch_resolving",
          "no_instruction", "math_pipe_throttle", "not_selected", "selected", "membar", "sleeping", "tex_throttle", "drain"]


def main():
    rep, sym, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {os.path.basename(rep)}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"\n## {d.get('Kernel Name')}  (launch id {d.get('ID')})")
        for k in KEYS:
            if k in d:
                print(f"{k:64s} {d[k]} {u.get(k, '')}")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in d:
                st.append((float(d[k] or 0), s))
        print("stalled warps per issue-active cycle: " + ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True)[:6]))
    src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
    srows = [r for r in csv.reader(io.StringIO(src)) if r and r[0].startswith("0x")]     # every launch repeats two header rows
    tmp = "/tmp/ncu_summary_cub"
    subprocess.call(["rm", "-rf", tmp])
    os.makedirs(tmp)
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "searcharray_b200", "libsearcharray_b200.so")],
                          cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dis = subprocess.check_output(["nvdisasm", "-g", cubin], cwd=tmp, stderr=subprocess.DEVNULL).decode().split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(f".text.{sym}:")][0]
    end = [i for i, l in enumerate(dis) if i > start and l.startswith("//--------------------- .text.")]
    end = end[0] if end else len(dis)
    cur, insts = None, []
    for l in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            insts.append(cur)
    n_per = len(insts)
    if n_per == 0 or len(srows) % n_per:
        print(f"\n(source attribution skipped: {len(srows)} sampled SASS rows vs {n_per} instructions in the shipped cubin)")
        return
    agg = collections.Counter()
    for i, r in enumerate(srows):
        agg[insts[i % n_per]] += int(r[2] or 0)
    tot = sum(agg.values()) or 1
    files = {}
    print(f"\n## source lines by warp-stall samples ({tot} samples)")
    for (f, ln), v in agg.most_common(14):
        if f not in files:
            p = os.path.join(ROOT, "searcharray_b200", "csrc", f)
            files[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = files[f][ln - 1].strip()[:100] if ln - 1 < len(files[f]) else ""
        print(f"{100 * v / tot:5.1f}%  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
