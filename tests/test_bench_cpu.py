"""bench.py's CPU-side contract (no GPU needed): the --impl reference arm prints ONE JSON line carrying the keys the driver
reads, on the SAME `config` object as the b200 arm (the driver compares the two), and rank != 0 stays silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra,
                       capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_line_and_config_match():
    import bench
    out = _run(["--config", "cfg1"])
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fft_gflops_5nlog2n" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config object as the b200 arm builds for this workload
    assert d["config"] == bench.config_object("cfg1", 1, "strong")
    assert d["scaling"] == bench.workload_geometry("cfg1", 1, "strong")[2]


def test_config_object_follows_the_rank_count():
    import bench
    for world in (1, 2, 4, 8):
        c = bench.config_object("cfg2", world, "strong")
        assert c["per_gpu_shape"] == [65536 // world, 4096]            # BASELINE's "2 @ P" split
        assert "%d of the 65536 rows" % (65536 // world) in c["sharding"]
        replicas, my_batch, scaling, shape, nbytes, nbuf = bench.workload_geometry("cfg2", world, "strong")
        assert (replicas, scaling, nbytes) == (1, "strong", (65536 // world) * 4096 * 16)
    # cfg1 fits L2: several rotating buffer pairs so every step streams from HBM
    assert bench.workload_geometry("cfg1", 1, "strong")[5] >= 4
    # single-GPU configs at N>1 are replicas
    assert bench.config_object("cfg3", 4, "strong")["sharding"] == "independent replicas"


def test_reference_arm_other_ranks_print_nothing():
    assert _run(["--config", "cfg1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""
