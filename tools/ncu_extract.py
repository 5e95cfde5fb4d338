"""Summarise an .ncu-rep (ncu --set full) as one CSV row per launch with the counters DESIGN.md / profiles cite.
   python tools/ncu_extract.py report.ncu-rep > summary.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(["id", "kernel"] + [n for m, n in WANT if m in col])
    for r in rows[2:]:
        vals = []
        for m, n in WANT:
            if m not in col:
                continue
            v, u = r[col[m]].replace(",", ""), units[col[m]]
            try:
                f = float(v)
                if n == "time_us":
                    f = f / 1e3 if u in ("ns", "nsecond") else (f * 1e3 if u in ("ms", "msecond") else f)
                if n.endswith("_MB"):
                    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * f
                vals.append(f"{f:.2f}")
            except ValueError:
                vals.append(v)
        out.writerow([r[col["ID"]], r[col["Kernel Name"]][:70]] + vals)


if __name__ == "__main__":
    main()
