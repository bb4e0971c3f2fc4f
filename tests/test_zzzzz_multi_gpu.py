"""GPU (-m gpu), needs >= 2 devices: the sharded step on REAL ranks (one process per GPU, NCCL inside the library) returns exactly
the single-GPU rows.  Skipped on one-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_zzzzz_multi_gpu.py -m gpu`."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_rows_equal_single_gpu_rows(world):
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--gpus", str(world), "--mbp", "5", "--steps", "2",
                        "--port", str(29600 + world)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["same_as_single_gpu"] and line["e2e_same"] and line["rows"] == line["single_rows"] > 10000
    # the sorted mer list sharded by mer range over the same ranks: the mer sequence of the single-GPU list, every position once
    assert line["sml_sharded"]["same_mer_sequence"] and line["sml_sharded"]["a_permutation"]
