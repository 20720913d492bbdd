"""Summarise an ncu report (raw + source page CSVs) the way profiles/*.md quote it."""
import csv, collections, sys
raw, src = sys.argv[1], sys.argv[2]
r = list(csv.reader(open(raw)))
d = {h: (u, v) for h, u, v in zip(r[0], r[1], r[2])}
def g(k):
    return d[k][1] if k in d else "n/a"
_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
def nbytes(k):
    """metric value in bytes: ncu auto-scales the unit per column (byte / Kbyte / Mbyte / Gbyte)"""
    if k not in d:
        return float("nan")
    u, v = d[k]
    return float(v.replace(",", "")) * _SCALE.get(u.strip(), 1.0)
print("duration_us", g("gpu__time_duration.sum"), "| cycles", g("sm__cycles_elapsed.max"), "| regs", g("launch__registers_per_thread"))
print("dram read MB %.3f write MB %.3f" % (nbytes("dram__bytes_read.sum") / 1e6, nbytes("dram__bytes_write.sum") / 1e6),
      "| dram %", g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
if len(sys.argv) > 3:   # optional: write the traffic record bench.py reads (roofline.traffic)
    import json
    json.dump({"kernel": r[2][r[0].index("Kernel Name")] if "Kernel Name" in r[0] else "?", "capture": raw,
               "dram_bytes_read": int(nbytes("dram__bytes_read.sum")), "dram_bytes_write": int(nbytes("dram__bytes_write.sum")),
               "dram_bytes_total": int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"))},
              open(sys.argv[3], "w"), indent=1)
print("warp inst", g("smsp__inst_executed.sum"), "| issue active %", g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "| warps active %", g("sm__warps_active.avg.pct_of_peak_sustained_active"))
print("pipes % : fma", g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), "alu", g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
      "lsu", g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), "xu", g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"))
print("shared wavefronts", g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), "ld", g("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"), "st", g("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
      "| conflicts", g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"))
for k in d:
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        v = float(d[k][1])
        if v > 0.05: print("  stall", k.split("issue_stalled_")[1].split("_per_issue")[0], round(v, 3))
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
ti = sum(f(r, "Instructions Executed") for r in data); ts = sum(f(r, "# Samples") for r in data)
mix = collections.Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] += f(r, "Instructions Executed")
print("opcode mix %:", ", ".join(f"{k} {v/ti*100:.1f}" for k, v in mix.most_common(14)))
bar = [i for i, r in enumerate(data) if "BAR.SYNC" in r[ix["Source"]]]
prev = 0
for b in bar + [len(data) - 1]:
    seg = data[prev:b + 1]
    print(f"  segment SASS[{prev}:{b}] inst {sum(f(r,'Instructions Executed') for r in seg)/ti*100:5.1f}%  samples {sum(f(r,'# Samples') for r in seg)/ts*100:5.1f}%  shared wf {sum(f(r,'L1 Wavefronts Shared') for r in seg)/1e6:6.1f}M (excess {sum(f(r,'L1 Wavefronts Shared Excessive') for r in seg)/1e6:5.1f}M)")
    prev = b + 1
