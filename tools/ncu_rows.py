"""Print one line per kernel launch of an `ncu --csv --metrics ...` log: time, DRAM bytes, L2 hit rate, instructions."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        data.setdefault((int(d["ID"]), d["Kernel Name"][:26]), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
tot = 0.0
for (i, name), v in data.items():
    t = v.get("gpu__time_duration.sum", 0) / 1e6
    tot += t
    print("%2d %-26s %7.3f ms  dram r %6.3f w %6.3f GB  L2 hit %5.1f%%  inst %6.1f M  sm %4.1f%% l1 %4.1f%% lts %4.1f%%" % (
        i, name, t, v.get("dram__bytes_read.sum", 0) / 1e9, v.get("dram__bytes_write.sum", 0) / 1e9,
        v.get("lts__t_sector_hit_rate.pct", 0), v.get("smsp__inst_executed.sum", 0) / 1e6,
        v.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0), v.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 0),
        v.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0)))
print("sum %.3f ms" % tot)
