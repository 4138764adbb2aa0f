"""Pull the handful of `ncu --set full` metrics the profile summaries quote out of a .ncu-rep (read here with `ncu -i ... --page raw --csv`).
    python tools/ncu_summary.py gpurun_out/r02_conv_2d128.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.max", "sm__pipe_tensor_op_hmma_cycles_active.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        print("==", path)
        name_i = hdr.index("Kernel Name")
        cols = [(h, i) for i, h in enumerate(hdr) if any(h == w or h.startswith(w) for w in WANT) or "tensor" in h and "cycles_active" in h]
        for r in data:
            print("--", r[name_i][:60])
            for h, i in cols:
                if r[i] not in ("", "n/a"):
                    print("   %-75s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
