"""Turn gpurun_out/ ncu artefacts of round 2 into the tracked summaries under profiles/:
    python tools/profile_r02.py <tag> <workload> <launches.csv> <full.ncu-rep>
  profiles/<tag>_<workload>_launch_shares.txt   per-kernel share of the frame (ncu --metrics gpu__time_duration.sum launch list)
  profiles/<tag>_<workload>_ncu_full_summary.txt the --set full numbers DESIGN.md quotes
  profiles/r02_traffic.json                      measured DRAM bytes per launch of the traversal kernels, stamped with the sha of the
                                                 kernel sources they were captured from (bench.py ignores it when the sources changed)."""
import collections, csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
tag, workload, launches, rep = sys.argv[1:5]
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
with open("profiles/%s_%s_launch_shares.txt" % (tag, workload), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python tools/run_workload.py %s 1 (one warm-up frame + one frame)\n" % workload)
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_share_of_step, not absolutes\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-50s launches=%4d total=%9.3f ms share=%.3f avg=%.3f ms\n" % (k, v[0], v[1], v[1] / tot, v[1] / v[0]))
summary = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
with open("profiles/%s_%s_ncu_full_summary.txt" % (tag, workload), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, python tools/run_workload.py %s 1, a few launches of closest / shade / shadow_any\n" % workload)
    f.write(summary)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}; units = rows[1]
def val(r, key):
    x = float(r[ix[key]].replace(",", "")); u = units[ix[key]]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
tscale = {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}
traffic = collections.defaultdict(list)
for r in rows[2:]:
    kn = r[ix["Kernel Name"]]
    name = "closest_kernel" if "closest_kernel" in kn else "shadow_any_kernel" if "shadow_any" in kn else None
    if name and float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * tscale.get(units[ix["gpu__time_duration.sum"]], 1.0) >= 50.0:
        traffic[name].append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))      # tail waves of a few rays (< 50 us) are not representative
path = "profiles/r02_traffic.json"
db = json.load(open(path)) if os.path.exists(path) else {}
if db.get("source_sha") != bench.source_sha():
    db = {"source_sha": bench.source_sha()}
dom = max(traffic, key=lambda k: sum(traffic[k])) if traffic else None
db[workload] = {"kernels": {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_profiled": len(v)} for k, v in traffic.items()},
                "source": "ncu --set full, profiles/%s_%s_ncu_full_summary.txt" % (tag, workload)}
json.dump(db, open(path, "w"), indent=1)
print(json.dumps(db[workload]))
