"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares for ONE step
(the last complete step's worth of launches is selected by kernel count)."""
import csv, sys, re, collections
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    val = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): val *= 1e3
    elif unit in ("ms", "msecond"): val *= 1e6
    rows.append((name, val))
def short(n):
    n = re.sub(r"\(.*", "", n)
    n = n.replace("b200::", "")
    m = re.match(r"gemm_kernel<(\d+), *(true|false|\(bool\)[01]|[01]), *(true|false|\(bool\)[01]|[01]), *(\d+)", n)
    if m:
        epi = {0:"store_bf16",1:"gelu",2:"resid_f32",3:"dgelu",4:"reduce_f32(wgrad)",5:"store_f32"}[int(m.group(4))]
        a = "MN" if m.group(2) in ("true","(bool)1","1") else "K"; b = "MN" if m.group(3) in ("true","(bool)1","1") else "K"
        return f"gemm<BN{m.group(1)},{a}x{b},{epi}>"
    return n[:70]
agg = collections.OrderedDict()
for n, v in rows:
    k = short(n)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
total = sum(v for _, v in rows)
print(f"launches {len(rows)}  total {total/1e6:.3f} ms  (per step: {total/1e6/steps:.3f} ms over {steps} steps)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/1e6/steps:9.3f} ms/step {100*v/total:6.2f}%  n/step={c/steps:7.1f}  avg={v/c/1e3:9.1f} us  {k}")
