"""Summarise .ncu-rep captures (read here, no GPU needed) into small CSV files under profiles/.

    python tools_dev/ncu_summary.py gpurun_out/foo.ncu-rep profiles/r01_foo
writes profiles/r01_foo_metrics.csv (selected raw metrics per captured launch) and profiles/r01_foo_stalls.csv (top source lines by
warp-stall samples)."""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out + "_metrics.csv", "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        ik = hdr.index("Kernel Name")
        for i, r in enumerate(data):
            for h, u, v in zip(hdr, units, r):
                name = h.split("TriageCompute.")[-1]
                if name in KEEP:
                    w.writerow([i, r[ik][:60], name, u, v])
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    hidx = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    nxt = next((i for i in range(hidx + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))     # first captured launch only
    hdr, data = rows[hidx], [r for r in rows[hidx + 1:nxt] if len(r) == len(rows[hidx]) and r[0] != "Address"]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[isamp] or 0) for r in data)
    with open(out + "_stalls.csv", "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow(["samples", "share", "executed", "sass", "top_stalls"])
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:25]:
            st = sorted(((hdr[i], int(r[i] or 0)) for i in stall if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:3]
            w.writerow([r[isamp], "%.3f" % (int(r[isamp] or 0) / max(total, 1)), r[iex], r[isrc][:100], " ".join("%s=%d" % kv for kv in st)])
    print(out, "stall samples:", total)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
