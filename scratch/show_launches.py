"""Prints one line per kernel launch from a scratch/launch_list.sh CSV."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
acc = collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        acc.setdefault((d["ID"], d["Kernel Name"][:48]), {})[d["Metric Name"]] = d["Metric Value"]
for (i, k), m in acc.items():
    f = lambda n: float(m.get(n, "nan").replace(",", ""))
    print(f"{i:>3} {k:48s} {f('gpu__time_duration.sum')/1e3:8.1f} us  inst {f('smsp__inst_executed.sum')/1e6:7.2f} M  occ {f('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} %  "
          f"issue {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %  dram r/w {f('dram__bytes_read.sum')/1e6:6.1f}/{f('dram__bytes_write.sum')/1e6:6.1f} MB  red sectors {f('lts__t_sectors_srcunit_tex_op_red.sum')/1e6:6.2f} M")
