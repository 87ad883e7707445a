"""Summarise an .ncu-rep (read on the CPU box with `ncu -i ... --page raw --csv`) into the handful of counters the
roofline argument needs.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit", "sm__inst_executed.sum", "smsp__average_warp_latency_issue_stalled", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe", "nvlrx__bytes.sum", "nvltx__bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__pipe_tensor_op"]


def main(path, header=None):
    if header:
        print(f"# {header}")
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d.get('Kernel Name', '?')[:160]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for h, u, v in zip(hdr, units, r):
            if any(h.endswith(k) or (k.endswith('_') and k in h) for k in KEYS) and v not in ("", "n/a"):
                print(f"   {h} = {v} {u}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
