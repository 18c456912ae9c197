"""Integration tier: BASELINE config #1 -- 2-layer MLP, 2 workers, real OS
processes on localhost ports via ``demo.py`` (how the reference is exercised
by hand)."""
import json
import os
import subprocess
import sys
import time
import urllib.request

import pytest

from fedtest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _get(url, timeout=5):
    with urllib.request.urlopen(url, timeout=timeout) as r:
        return r.status, json.loads(r.read().decode())


def _wait(pred, timeout=60, step=0.1):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            if pred():
                return True
        except Exception:
            pass
        time.sleep(step)
    return False


@pytest.mark.slow
def test_mlp2_two_workers_subprocess_round_trip():
    mport, w1, w2 = free_port(), free_port(), free_port()
    env = dict(os.environ, PYTHONPATH=ROOT, PYTHONUNBUFFERED="1")
    common = ["--model", "mlp2", "--lr", "0.01", "--bind", "127.0.0.1", "--heartbeat-time", "1"]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "demo.py"), "manager", "x", str(mport)] + common,
                              env=env, cwd=ROOT)]
    base = "http://127.0.0.1:{}/mlp2/".format(mport)
    try:
        assert _wait(lambda: _get(base + "clients")[0] == 200), "manager did not come up"
        for p in (w1, w2):
            procs.append(subprocess.Popen(
                [sys.executable, os.path.join(ROOT, "demo.py"), "worker", "127.0.0.1:{}".format(mport), str(p)] + common,
                env=env, cwd=ROOT))
        assert _wait(lambda: len(_get(base + "clients")[1]) == 2), "workers did not register"
        for rnd in range(3):
            status, body = _get(base + "start_round?n_epoch=3", timeout=30)
            assert status == 200 and len(body) == 2 and all(body.values())
            assert _wait(lambda: _get(base + "state")[1]["n_updates"] == rnd + 1), "round did not finish"
        status, hist = _get(base + "loss_history")
        assert status == 200 and len(hist) == 9 and hist[-1] < hist[0]
        clients = _get(base + "clients")[1]
        assert all(c["num_updates"] == 3 for c in clients)
    finally:
        for p in procs:
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
