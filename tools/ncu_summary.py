"""Summarise an .ncu-rep: headline metrics per kernel + stall samples by opcode / top source lines.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] [--top N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if filt and filt not in name:
        continue
    print("==", name[:110])
    for w in want:
        if w in h:
            print("   %-70s %s %s" % (w, r[h.index(w)], rows[1][h.index(w)]))
    st = [(c, float(r[i])) for i, c in enumerate(h) if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio") and r[i]]
    st.sort(key=lambda kv: -kv[1])
    print("   stalls/issue:", ", ".join("%s %.2f" % (c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for c, v in st[:7]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = None
blocks = []
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        kern = r[1]
    if r and r[0] == "Address":
        blocks.append((kern, i))
for bi, (kern, i) in enumerate(blocks):
    if filt and filt not in (kern or ""):
        continue
    hdr = rows[i]
    end = len(rows)
    for j in range(i + 1, len(rows)):
        if rows[j] and rows[j][0] == "Kernel Name":
            end = j
            break
    body = [r for r in rows[i + 1:end] if len(r) == len(hdr)]
    si, so = hdr.index("# Samples"), hdr.index("Source")
    stall_cols = [k for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in body) or 1
    agg = collections.defaultdict(int)
    for r in body:
        t = r[so].split()
        op = (t[1] if t and t[0].startswith("@") else t[0] if t else "?").split(".")[0]
        agg[op] += int(r[si])
    print("== source:", (kern or "")[:100], "samples", tot)
    print("   by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
    for r in sorted(body, key=lambda r: -int(r[si]))[:top_n]:
        reasons = sorted(((hdr[k], int(r[k])) for k in stall_cols if int(r[k]) > 0), key=lambda kv: -kv[1])[:2]
        print("   %5.1f%%  %-70s %s" % (100.0 * int(r[si]) / tot, r[so].strip()[:70], reasons))
    break
