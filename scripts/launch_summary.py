"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ix = {h: i for i, h in enumerate(hdr)}
c = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    c.setdefault(r[ix["Kernel Name"]][:70], []).append(v)
tot = sum(sum(v) for v in c.values())
print("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, v in c.items():
    print(f"| `{k}` | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | {sum(v) / tot:.3f} |")
