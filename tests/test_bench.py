"""bench.py on the CPU: the reference arm (the reference's algorithm on the host cores) must run without a GPU,
without mapping the product library into its process, and print the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_on_cpu_without_the_product_library(oracle):
    code = (
        "import sys, runpy\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-traj-per-core', '1']\n"
        "try:\n"
        "    runpy.run_path(%r, run_name='__main__')\n"
        "except SystemExit as e:\n"
        "    assert not e.code, e.code\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('MAPPED', 'librebop_b200' in maps, 'rebop_b200' in sys.modules, file=sys.stderr)\n"
    ) % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MAPPED False False" in r.stderr, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ssa_reaction_events_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    # the workload is the GPU arm's; the bounded sample that was actually timed is stated next to it
    assert "1250000 trajectories per GPU" in line["config"]["workload"]
    assert line["trajectories_per_step"] == line["cpu_baseline"]["cores"] and str(line["trajectories_per_step"]) in line["sample"]


def test_gpu_arm_fails_loudly_without_a_device():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
