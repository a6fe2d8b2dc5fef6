"""Print selected metrics from an `ncu --page raw --csv` export.   python tools/ncu_keys.py file.csv [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]; data = rows[2:]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
  r"gpu__time_duration.sum|sm__pipe_tensor.*cycles_active.*pct|sm__inst_executed_pipe_(fp64|uniform|lsu).*pct|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|lts__t_sector_hit_rate.pct|lts__t_bytes.sum$|lts__throughput.avg.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__issue_active.avg.pct|sm__throughput.avg.pct|l1tex__throughput.avg.pct|lts__t_sectors_srcunit_tex_op_read.sum$|sm__cycles_elapsed.max|smsp__cycles_active.avg|l1tex__m_xbar2l1tex_read_bytes.sum$|sm__pipe_fp64_cycles_active.*pct|smsp__warp_issue_stalled.*_per_warp_active.pct")
for d in data:
    print("##", d[hdr.index("Kernel Name")][:60])
    for h, u, v in zip(hdr, units, d):
        if pat.search(h): print("  %-90s %s %s" % (h, v, u))
