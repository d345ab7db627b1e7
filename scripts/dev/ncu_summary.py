"""Summarise an `ncu --set full` report (read here with `ncu -i`): per-kernel JSON of the metrics DESIGN.md / bench.py cite,
plus the DRAM-traffic file bench.py reads.  usage: ncu_summary.py <report.ncu-rep> <out.json> [traffic.json "<source note>"]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
t = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
t = t[t.index('"ID"'):]
rows = list(csv.reader(io.StringIO(t)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum")
res = []
for r in data:
    d = {}
    for h, u, v in zip(hdr, units, r):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try: d[h] = float(v.replace(",", ""))
            except ValueError: d[h] = v
            if u and h.startswith("dram__bytes"): d[h + ".unit"] = u
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
for d in res:
    print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), "rd", d.get("dram__bytes_read.sum"), d.get("dram__bytes_read.sum.unit"),
          "wr", d.get("dram__bytes_write.sum"), "issue%", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "l1tex%", d.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), "regs", d.get("launch__registers_per_thread"))
if len(sys.argv) > 3:
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for d in res:
        b = sum(d[k] * mult[d[k + ".unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per[d["Kernel Name"][:60]] = b
    json.dump({"dram_bytes_per_matvec_B16_f32": sum(per.values()), "per_kernel": per, "source": sys.argv[4],
               "algorithmic_bytes_per_matvec_B16_f32": 135992000}, open(sys.argv[3], "w"), indent=1)
    print("traffic", sum(per.values()))
