"""Digest of an .ncu-rep: headline metrics, warp-stall breakdown, and instructions / stall samples per SOURCE LINE
(the SASS page of the report joined with `nvdisasm -g` line info of the same object; build with -lineinfo).
Usage: python tools/ncu_digest.py file.ncu-rep [object.o kernel_name_substring [top_n [source_root]]]
(object.o and source_root must be the build the capture was taken from: for an older capture, check that commit out into
a scratch worktree, compile the one .cu with -lineinfo and pass both)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else None
kname = sys.argv[3] if len(sys.argv) > 3 else None
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
src_root = sys.argv[5] if len(sys.argv) > 5 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    ud = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "?")[:120])
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {ud[k]}")
    st = []
    for h, v in d.items():
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(v.replace(",", "")), h))
            except ValueError:
                pass
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {v:.3f}")

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {c: i for i, c in enumerate(h)}
insts = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[0].startswith("0x"):
        continue
    insts.append((int(r[0], 16), r[ci["Source"]].strip(), float(r[ci["# Samples"]] or 0), float(r[ci["Instructions Executed"]] or 0),
                  {k: float(r[ci[k]] or 0) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_no_inst", "stall_math")}))
base = insts[0][0]
tot_s = sum(i[2] for i in insts)
tot_i = sum(i[3] for i in insts)
print(f"SASS instructions {len(insts)}, warp instructions executed {tot_i:.4g}, samples {tot_s:.0f}")
if obj and kname:
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")]
        dis = ""
        for c in cub:
            dis += subprocess.run(["nvdisasm", "-g", os.path.join(td, c)], capture_output=True, text=True).stdout
    line_of = {}
    cur, on = None, False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = kname in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    for addr, text, s, ie, stl in insts:
        key = line_of.get(addr - base, ("?", 0))
        a = agg[key]
        a[0] += s
        a[1] += ie
        for k, v in stl.items():
            a[2][k] += v
    srcs = {}
    print("samples  share | warp-instr share | file:line | top stalls | source")
    for key, (s, ie, stl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        f, l = key
        if f not in srcs:
            p = next((os.path.join(dp, f) for dp, _, fs in os.walk(src_root) if f in fs), None)
            srcs[f] = open(p).read().splitlines() if p else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        top = ",".join(f"{k.replace('stall_', '')}:{v / max(s, 1):.2f}" for k, v in stl.most_common(2))
        print(f"{s:8.0f} {100 * s / tot_s:5.1f}% | {ie:10.4g} {100 * ie / tot_i:5.1f}% | {f}:{l} | {top} | {text}")
else:
    for addr, text, s, ie, stl in sorted(insts, key=lambda x: -x[2])[:topn]:
        print(f"{s:8.0f} {100 * s / tot_s:5.1f}% {ie:12.4g}  +{addr - base:05x}  {text}")
