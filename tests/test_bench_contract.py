"""bench.py contract, CPU part: the reference arm prints exactly ONE JSON line on stdout with the keys the driver reads
(the GPU arm shares the code path that formats the line; it is exercised on the B200 by tools/gpu_round.sh)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec at 4096 AntGather envs/GPU" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--skip-cpu"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0 and res.stdout.strip() == ""   # no CPU fallback, nothing that looks like a result


def test_real_reference_arm_runs_against_a_fake_pybullet_stack(tmp_path):
    """bench/ref_pybullet_mp.py (the UNMODIFIED reference in a multiprocessing vector env) must work the day pybullet is
    importable.  Here `pybullet`, `gym` and `hrl_pybullet_envs` are tiny fakes on PYTHONPATH (the worker processes are
    spawned, so they import them like the real ones): the arm reports kind "reference" through bench.py's own code path."""
    (tmp_path / "pybullet.py").write_text("")
    gym = tmp_path / "gym"; gym.mkdir()
    (gym / "__init__.py").write_text(
        "import numpy as np\n"
        "class _Space:\n    shape = (8,)\n"
        "class _Env:\n"
        "    action_space = _Space()\n"
        "    def __init__(self): self.t = 0\n"
        "    def seed(self, s): return [s]\n"
        "    def reset(self): self.t = 0; return np.zeros(46)\n"
        "    def step(self, a):\n"
        "        assert np.shape(a) == (8,)\n"
        "        self.t += 1\n"
        "        return np.zeros(46), 1.0, self.t % 7 == 0, {}\n"
        "def make(env_id):\n    assert env_id == 'AntGatherBulletEnv-v0'; return _Env()\n")
    ref = tmp_path / "hrl_pybullet_envs"; ref.mkdir()
    (ref / "__init__.py").write_text("")
    env = dict(os.environ, PYTHONPATH=str(tmp_path) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "30", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([l for l in res.stdout.split("\n") if l.strip()][0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["cores"] >= 1 and "worker" in d["config"]["note"]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench", "ref_pybullet_mp.py"), "--steps", "20", "--warmup", "2", "--workers", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    r = json.loads(out.stdout.strip().split("\n")[-1])
    assert r["workers"] == 2 and r["episodes"] >= 2 and r["mean_return"] == 7.0   # episodes of 7 steps, reward 1 each
    # and without the fakes the same command says so
    plain = subprocess.run([sys.executable, os.path.join(ROOT, "bench", "ref_pybullet_mp.py")], capture_output=True, text=True, timeout=120, cwd=ROOT)
    import importlib.util
    if importlib.util.find_spec("pybullet") is None:
        assert "reference unavailable: pybullet not installed" in plain.stdout
