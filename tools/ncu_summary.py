"""Summarise an .ncu-rep (raw page CSV) into the handful of metrics the roofline discussion needs."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = {
 "Kernel Name": "kernel", "gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
 "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active": "hmma_pct",
 "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
 "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
 "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct", "sm__cycles_elapsed.avg": "cycles",
 "smsp__inst_executed.sum": "inst", "lts__t_bytes.sum": "l2_bytes", "sm__cycles_elapsed.avg.per_second": "sm_hz",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conflicts", "launch__grid_size": "grid",
 "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
}
idx = {hdr.index(k): v for k, v in want.items() if k in hdr}
res = []
for r in rows[2:]:
    d = {}
    for i, name in idx.items():
        v = r[i]
        try: v = float(v.replace(",", ""))
        except Exception: v = v[:70]
        d[name] = v; 
        if name != "kernel": d[name + "_unit"] = units[i]
    res.append(d)
for d in res:
    print(d["kernel"])
    print("   ", {k: v for k, v in d.items() if k != "kernel" and not k.endswith("_unit")})
    print("   units:", {k[:-5]: v for k, v in d.items() if k.endswith("_unit") and k[:-5] in ("time","dram_rd","dram_wr","l2_bytes","sm_hz")})
