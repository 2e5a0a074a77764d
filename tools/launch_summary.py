"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv) into profiles/: keeps the rows of this
library's kernels (namespace cmf) and appends a per-kernel table of count / mean time / share of a step.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_launches.csv
"""
import collections
import csv
import io
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src).read()
    start = text.index('"ID"')
    rows = list(csv.reader(io.StringIO(text[start:])))
    hdr, data = rows[0], rows[1:]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    mine = [r for r in data if len(r) > iv and "cmf::" in r[ik]]
    agg = collections.OrderedDict()
    for r in mine:
        name = r[ik].split("(")[0].replace("void ", "").replace("cmf::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write("# per-kernel summary of %d launches of this library (times under ncu: cold cache, serialised)\n" % len(mine))
        fh.write("# kernel,launches,mean_us,share_of_total,grid,block\n")
        for name, (cnt, ns, grid, block) in agg.items():
            fh.write("# %s,%d,%.1f,%.4f,%s,%s\n" % (name, cnt, ns / cnt / 1e3, ns / total, grid.replace(",", " "), block.replace(",", " ")))
        w = csv.writer(fh)
        w.writerow([hdr[0], hdr[ik], hdr[ib], hdr[ig], "gpu__time_duration.sum [ns]"])
        for r in mine:
            w.writerow([r[0], r[ik].split("(")[0], r[ib], r[ig], r[iv]])
    print(open(dst).read().split('"ID"')[0] if False else "".join(l for l in open(dst) if l.startswith("#")))


if __name__ == "__main__":
    main()
