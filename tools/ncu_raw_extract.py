"""Pick the metrics the roofline discussion needs out of an `ncu --page raw --csv` export: one block per launch."""
import csv, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum",
]

def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {n: i for i, n in enumerate(hdr)}
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        print(f"== {r[col['Kernel Name']][:90]}  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}")
        for w in WANT:
            if w in col:
                print(f"   {w:75s} {r[col[w]]:>18s} {units[col[w]]}")
        rd, wr, du = (col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum"), col.get("gpu__time_duration.sum"))
        if rd and wr and du:
            def val(i):
                v = float(r[i].replace(",", ""))
                u = units[i]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "second": 1,
                            "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1)
            tb, t = val(rd) + val(wr), val(du)
            print(f"   {'dram traffic / duration':75s} {tb/1e6:14.2f} MB  {t*1e6:10.1f} us  {tb/t/1e9:8.1f} GB/s")

if __name__ == "__main__":
    main(sys.argv[1])
