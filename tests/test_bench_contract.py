"""CPU tests of bench.py's host-side bookkeeping: the per-op roofline table built from a recorded op table of a real
B200 run, the rank sharding of the retrieval job, and the JSON line of the reference arm (`--impl reference` runs the
oracle port on the host cores: the one place outside tests/ that may execute oracle/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_op_of_a_recorded_step_has_an_algorithmic_cost_model():
    """profiles/op_table_r3t.json = the C-ABI calls of one benchmark step (B200, final code of round 2).  Every tagged
    call must map to SURVEY 8(d) algorithmic bytes / flops and a bound; the hot ops must sit on the roof DESIGN names."""
    import bench
    with open(os.path.join(ROOT, "profiles", "op_table_r3t.json")) as f:
        table = json.load(f)
    per_step = {k: (v["ms_per_step"], v["calls_per_step"]) for k, v in table.items()}
    peaks = {"hbm_gbs": 6542.7, "bf16_tflops": 1623.4, "bf16_tflops_sustained": 1378.4, "source": "test"}
    rows, step_bytes = bench.op_roofline_table(per_step, peaks)
    assert len(rows) == len(per_step)
    by_op = {r["op"]: r for r in rows}
    tagged = [r for r in rows if "[" in r["op"]]
    assert tagged and all(r.get("bound") in ("hbm", "tensor", "alu", "latency") for r in tagged), \
        [r["op"] for r in tagged if r.get("bound") is None]
    assert all(r["algorithmic_mb"] > 0 and 0 < r["frac_hbm"] < 1.0 for r in tagged)
    head = by_op["dh3d_linear_rowdot_packed[M262144_K256_N1024]"]
    assert head["bound"] == "tensor" and head["calls"] == 2 and 0.2 < head["frac_bf16_burst"] < 1.0
    assert by_op["dh3d_flex_conv_pm_packed[n262144_K8_Ci64_Co64]"]["bound"] == "hbm"
    assert by_op["dh3d_netvlad[B32_N8192_D256_Kc64_O256]"]["bound"] == "hbm"
    # 32 clouds x ~97.6 MB of per-op compulsory traffic
    assert 2.5e9 < step_bytes < 3.5e9


def test_retrieval_shards_cover_the_job_once():
    from dh3d_b200.dist import shard_range
    for total, world in ((4096, 8), (4096, 1), (100, 3), (7, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_reference_arm_prints_the_contract_line():
    """One step of one cloud through the oracle port; the line carries the arm's own keys and no device copies."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "point_clouds_per_sec_full_dh3d_forward_n8192"
    assert d["unit"] == "clouds/s" and d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
