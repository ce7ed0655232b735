"""stdin: `ncu --page raw --csv` of one capture -> the few raw metrics the roofline needs (per launch)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        print("# raw metrics (launch %d):" % k)
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w:
                    print("#   %-60s %s %s" % (h, vals[i], units[i]))
