"""bench.py contract pieces that run without a GPU: the reference arm (CPU oracle only, libb2d.so never loaded), the
`config` object shared by both arms, host-binding helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_line(*extra):
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--poses", "6"] + list(extra),
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_reference_arm_line_and_config():
    d = _ref_line()
    assert d["impl"] == "reference" and d["metric"].startswith("frames/sec at 1920x1080") and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "u8" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "per core" in cb["sample"]
    # the config object is built by the function the GPU arm uses, from the same workload description
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    from oracle.host import OracleScene
    w, h, maps, n, desc = bench.workload("c2", argparse.Namespace(poses=6))
    sc = OracleScene(bench.build_wad(*maps[0][:3]), 0)
    assert d["config"] == bench.bench_config(desc, n, 1, sc.info)
    assert "configs[1]" in d["config"]["workload"] and d["config"]["poses_per_step_per_gpu"] == 6


def test_reference_arm_does_not_load_the_product_library():
    prog = """
import runpy, sys
sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--poses', '4']
try:
    runpy.run_path('bench.py', run_name='__main__')
except SystemExit:
    pass
maps = open('/proc/self/maps').read()
assert 'libb2d.so' not in maps, 'the reference arm mapped libb2d.so'
assert 'libb2d_oracle.so' in maps
print('clean')
"""
    res = subprocess.run([sys.executable, "-c", prog], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "clean" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_other_configs_describe_baseline_json():
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(poses=0)
    w, h, maps, n, desc = bench.workload("c3", ns)
    assert (w, h, len(maps), n) == (1920, 1080, 9, 1000) and [m[0] for m in maps] == ["E1M%d" % i for i in range(1, 10)]
    w, h, maps, n, desc = bench.workload("c4", ns)
    assert (w, h, len(maps), n) == (3840, 2160, 10, 1000) and maps[0][0] == "MAP01" and maps[-1][0] == "MAP10"
    w, h, maps, n, desc = bench.workload("c5", ns)
    assert (w, h, n) == (1920, 1080, 100000) and maps[0][3] == "random" and maps[0][4] == 5


def test_host_binding_helpers():
    from rust_doom_b200 import jobs
    assert jobs._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert 1 <= jobs.usable_cores() <= (os.cpu_count() or 1)
    assert jobs.gpu_numa_node(0) is None or isinstance(jobs.gpu_numa_node(0), int)
