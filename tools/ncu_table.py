"""One line per kernel launch of an .ncu-rep (--set full): time, DRAM bytes, DRAM / tensor-pipe / shared-operand utilisation,
registers, top stall reasons.  Usage: python tools/ncu_table.py rep.ncu-rep > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]


def col(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name):
            return i
    return None


K = {"name": col("Kernel Name"), "us": col("gpu__time_duration.sum"), "rd": col("dram__bytes_read.sum"), "wr": col("dram__bytes_write.sum"),
     "dram": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "tc": col("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
     "smem": col("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
     "issue": col("smsp__issue_active.avg.pct_of_peak_sustained_active"), "regs": col("launch__registers_per_thread"),
     "grid": col("launch__grid_size")}
units = rows[1]
stall_cols = [(i, h.split("stalled_")[1].split("_not_issued")[0].split(".")[0]) for i, h in enumerate(hdr)
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


def to_bytes(v, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return fnum(v) * m.get(unit, 1.0)


print("%-4s %-44s %9s %9s %9s %6s %6s %6s %6s %5s  %s" % ("#", "kernel", "time us", "dram rd MB", "dram wr MB", "dram%", "tensor%", "smemop%",
                                                           "issue%", "regs", "top stalls"))
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    name = r[K["name"]].replace("void ", "").replace("skb::", "")[:44]
    t = fnum(r[K["us"]]) * (1e-3 if units[K["us"]] == "ns" else (1e3 if units[K["us"]] == "ms" else 1.0))
    st = sorted(((fnum(r[i]), s) for i, s in stall_cols if r[i] not in ("", "n/a")), reverse=True)
    tot = sum(v for v, _ in st if v == v) or 1.0
    stalls = ", ".join("%s %.0f%%" % (s, 100 * v / tot) for v, s in st[:3])
    print("%-4d %-44s %9.1f %9.1f %9.1f %6.1f %6.1f %6.1f %6.1f %5s  %s" % (
        n, name, t, to_bytes(r[K["rd"]], units[K["rd"]]) / 1e6, to_bytes(r[K["wr"]], units[K["wr"]]) / 1e6, fnum(r[K["dram"]]),
        fnum(r[K["tc"]]), fnum(r[K["smem"]]), fnum(r[K["issue"]]), r[K["regs"]], stalls))
