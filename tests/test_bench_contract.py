"""CPU suite: bench.py's JSON-line contract.  The reference arm (`--impl reference`: the CPU oracle on the host cores) runs
for real on a bounded sample; the GPU arm's line is checked on the committed result of the last B200 run."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d) and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("Mrays/s") and d["unit"] == "Mrays/s" and d["vs_baseline"] is None and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "2x2 cell" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")))
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline", "workloads"} <= set(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "f32" and d["gpu_launches"] > 0 and d["scaling"] == "strong"
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "bounds"} <= set(rf) and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-6
    b = rf["bounds"]                                                # north_star: the lesser of bytes/ray over bandwidth and flops/ray over the FP32 peak
    assert b["binding"] in ("bandwidth", "fp32") and min(b["bound_bw_Grays_per_s"], b["bound_fp32_Grays_per_s"]) > 0 and b["l2_read_GBps_measured"] > 0
    assert set(d["workloads"]) == {"c4_standin", "c3_standin", "c5"}
    for w in d["workloads"].values():
        assert w["value"] > 0 and w["ms_per_frame"] > 0 and {"roofline", "cpu_baseline", "e2e", "config"} <= set(w) and "workload" in w["config"]
        assert "STAND-IN" in w["config"]["workload"] or "configs[4]" in w["config"]["workload"]
    for n in (2, 4, 8):                                             # the same frame at every N
        dn = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n%d.json" % n)))
        assert dn["n_gpus"] == n and dn["scaling"] == "strong" and dn["config"]["samples"] == d["config"]["samples"] and abs(dn["rays_per_step"] - d["rays_per_step"]) < 1e-6 * d["rays_per_step"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 1280 * 720 * 24 + d["e2e"]["d2h_bytes_per_step"] % (1280 * 720 * 24)
    assert d["e2e"]["value"] <= d["value"] * 1.001                  # host copies are inside the timed region
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
