"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (launch-list shares, full-set summary,
and profiles/r01_traffic.json = measured DRAM bytes per launch of the traversal kernels, read by bench.py)."""
import collections, csv, io, json, subprocess, sys
tag = sys.argv[1]
launches, rep = sys.argv[2], sys.argv[3]
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("rtx::", "")[:48]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open("profiles/%s_launch_shares.txt" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 1 --no-cpu-baseline`\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_share_of_step, not absolutes\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-50s launches=%4d total=%9.3f ms share=%.3f avg=%.3f ms\n" % (k, v[0], v[1], v[1] / tot, v[1] / v[0]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}; units = rows[1]
def val(r, key):
    x = float(r[ix[key]].replace(",", "")); u = units[ix[key]]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
traffic = collections.defaultdict(list)
for r in rows[2:]:
    name = "closest_kernel" if "closest_kernel" in r[ix["Kernel Name"]] else "shadow_any_kernel" if "shadow_any" in r[ix["Kernel Name"]] else None
    if name and float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1.0) >= 50.0:
        traffic[name].append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))          # tail waves of a few rays (< 50 us) are not representative
out = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_profiled": len(v), "source": "ncu --set full, profiles/%s_ncu_full_summary.txt" % tag} for k, v in traffic.items()}
json.dump(out, open("profiles/r01_traffic.json", "w"), indent=1)
print(json.dumps(out))
