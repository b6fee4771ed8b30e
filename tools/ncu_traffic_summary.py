"""Per-kernel duration, share and DRAM traffic of one NLL+gradient evaluation from an ncu CSV
(--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum); writes the JSON bench.py reads
for roofline.traffic.   python tools/ncu_traffic_summary.py profiles/r01_ncu_traffic_nll_n32768.csv 32768 > profiles/r01_traffic_n32768.json"""
import collections, csv, json, re, sys

path, n = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
for row in csv.DictReader(lines):
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").strip()
    v = float(row["Metric Value"].replace(",", ""))
    u, m = row["Metric Unit"], row["Metric Name"]
    if m == "gpu__time_duration.sum":
        agg[k]["ms"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    else:
        agg[k]["read_bytes" if "read" in m else "write_bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    ids[k].add(row["ID"])
tot = sum(a["ms"] for a in agg.values())
kern = {k: {"launches": len(ids[k]), "ms": round(a["ms"], 3), "share": round(a["ms"] / tot, 4), "dram_read_bytes": a["read_bytes"],
            "dram_write_bytes": a["write_bytes"]} for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
dmma = [k for k in kern if "potrf_ll" in k or "gemm_f64" in k]
out = {"matrix_order": n, "source": path, "total_ms_under_ncu": round(tot, 2), "kernels": kern,
       "dmma_kernels": dmma,
       "dmma_dram_bytes_per_evaluation": sum(kern[k]["dram_read_bytes"] + kern[k]["dram_write_bytes"] for k in dmma)}
print(json.dumps(out, indent=1))
