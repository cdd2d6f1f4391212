"""Turn an .ncu-rep into the text summaries kept under profiles/ (run in the build container):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name"""
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(det)))
    hdr = rows[0]
    ki, si, mi, ui, vi = (hdr.index(x) for x in ("Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value"))
    with open(out + "_details.txt", "w") as f:
        f.write("kernel: %s\n" % rows[1][ki])
        for r in rows[1:]:
            if r[mi]:
                f.write(f"{r[si][:34]:34s} {r[mi][:48]:48s} {r[ui]:16s} {r[vi]}\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
            "smsp__cycles_active.avg", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.min.pct_of_peak_sustained_active",
            "smsp__issue_active.max.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled")
    with open(out + "_raw.txt", "w") as f:
        for h, u, v in zip(rows[0], rows[1], rows[2]):
            if any(h == k or (k.startswith("smsp__average_warps_issue_stalled") and h.startswith(k)) for k in keep):
                f.write(f"{h:92s} {u:14s} {v}\n")


if __name__ == "__main__":
    main()
