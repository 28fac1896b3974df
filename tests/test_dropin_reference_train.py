"""Drop-in check: the REFERENCE's own, unmodified `train.py` (taken from the read-only reference checkout at test time into
pytest's temp dir, so that the script's directory does not put the reference package first on `sys.path`; nothing of it
lives in this repo) runs against this framework — its `import internlm...` lines resolve to `internevo_b200` through the alias package
— and trains the demo config on 2 CPU ranks to the same final loss as our `train.py`."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRAIN = "/root/reference/train.py"


def _losses(script, port, cwd):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), script, "--config", os.path.join(ROOT, "configs", "demo.py"),
                        "--launcher", "torch", "--backend", "gloo"], cwd=cwd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return [float(x) for x in re.findall(r"step=\d+ loss=([0-9.]+)", r.stdout + r.stderr)]


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="reference checkout not present")
def test_reference_train_py_runs_on_this_framework(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import find_free_port

    import shutil

    script = str(tmp_path / "reference_train.py")
    shutil.copy(REF_TRAIN, script)
    theirs = _losses(script, find_free_port(), str(tmp_path))
    ours = _losses(os.path.join(ROOT, "train.py"), find_free_port(), str(tmp_path))
    assert len(theirs) == len(ours) == 20
    assert theirs[-1] < 1.5 < theirs[0]
    assert all(abs(a - b) < 1e-4 for a, b in zip(theirs, ours)), (theirs, ours)
