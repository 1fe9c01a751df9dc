"""Condense an `ncu --set full --page raw --csv` export into the columns the roofline discussion needs."""
import csv
import re
import sys

WANT = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read_MB"),
        ("dram__bytes_write.sum", "dram_write_MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print("kernel," + ",".join(n for _, n in WANT))
    for r in rows[2:]:
        name = re.sub(r"\(CUtensorMap.*", "", r[ki]).replace("void ", "").replace("umma_gemm_kernel", "umma")
        vals = []
        for m, n in WANT:
            if m not in h:
                vals.append("")
                continue
            v, u = r[h.index(m)], units[h.index(m)]
            try:
                x = float(v.replace(",", ""))
                if n == "time_us" and u in ("ns", "nsecond"):
                    x /= 1e3
                if n == "time_us" and u in ("ms", "msecond"):
                    x *= 1e3
                if n.endswith("_MB") and u in ("byte", "Byte"):
                    x /= 1e6
                if n.endswith("_MB") and u in ("Kbyte",):
                    x /= 1e3
                if n.endswith("_MB") and u in ("Gbyte",):
                    x *= 1e3
                vals.append("%.2f" % x)
            except ValueError:
                vals.append(v)
        print('"%s",%s' % (name, ",".join(vals)))


if __name__ == "__main__":
    main(sys.argv[1])
