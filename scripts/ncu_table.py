"""Print per-launch rows of an ncu --csv log (last forward only with --last)."""
import csv, re, sys, collections
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", re.sub(r"sea::<unnamed>::", "", r["Kernel Name"]))[:46],
                                   "grid": r["Grid Size"], "blk": r["Block Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * (1e3 if r["Metric Unit"] in ("us",) else 1e6 if r["Metric Unit"] == "ms" else 1)
    d["unit:" + r["Metric Name"]] = r["Metric Unit"]
items = list(by_id.values())
if "--last" in sys.argv:
    idx = [i for i, d in enumerate(items) if "tipi_hidden" in d["name"]]
    items = items[idx[-1]:]
tot = 0
for d in items:
    t = d.get("gpu__time_duration.sum", 0.0)
    def b(k):
        v = d.get(k, 0.0); u = d.get("unit:" + k, "byte")
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
    tot += t
    print(f"{d['name']:48s} {d['grid']:>14s} {t/1e3:8.1f} us  rd {rd/1e6:7.1f} MB wr {wr/1e6:7.1f} MB  {(rd+wr)/max(t,1):6.2f} GB/s/1e0".replace("GB/s/1e0", "TB/s x1e-3") )
print(f"total {tot/1e3:.1f} us over {len(items)} launches")
