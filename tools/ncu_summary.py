"""Key metrics + top stall instructions from an .ncu-rep (needs ncu on PATH).  Usage: ncu_summary.py rep [launch_idx]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2 + idx]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]
for k in keys:
    for i, h in enumerate(hdr):
        if h == k or h.endswith("." + k):
            print("%-90s %-10s %s" % (h, units[i], r[i]))
            break
st = []
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st.append((float(r[i]), h.split("stalled_")[1]))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1
print("stalls:", ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
i_src, i_s = h.index("Source"), h.index("# Samples")
data = []
seen = set()
for k, row in enumerate(rows[2:]):
    try:
        key = (row[h.index("Address")])
        if key in seen:
            continue
        seen.add(key)
        data.append((int(row[i_s]), k, row[i_src]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print("top stall instructions (%d samples):" % tot)
ctx = int(__import__("os").environ.get("NCU_CTX", "0"))        # SASS lines of context before each hot instruction
for s, k, t in sorted(data, reverse=True)[:22]:
    for kk in range(max(0, k - ctx), k):
        print("                 | %s" % rows[2 + kk][i_src][:100])
    print("  %5.1f%%  #%4d  %s" % (100.0 * s / tot, k, t[:100]))
