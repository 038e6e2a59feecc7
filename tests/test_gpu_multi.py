"""GPU, REAL multi-rank: one process per GPU over NCCL (skipped with fewer than two visible GPUs).  The frame-range split of
one long signal (parallel.long_form_features: partial -> NCCL all-reduce of the statistics -> finalize) and the clip split
must reproduce the oracle's whole-signal result (parakeet/audio.py:39-78 with the whole-file statistics of :66-69;
whisper/audio.py:83's single max)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import wrappers_oracle as W  # noqa: E402
from oracle.make_golden import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(world, out_path, length):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_long_form.py"), out_path, str(length)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            tails = [(q.communicate()[0] or "")[-1500:] for q in procs]
            raise AssertionError("multi-GPU worker timed out:\n" + "\n-----\n".join(tails))
        logs.append(out)
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{logs[r][-3000:]}"


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_long_form_and_clip_sharding_on_real_gpus(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    length = 16000 * 60 + 137  # one minute and a ragged tail: frame counts differ between ranks
    out_path = str(tmp_path / "out.npz")
    _launch(world, out_path, length)
    got = np.load(out_path)
    assert int(got["world"]) == world and bytes(got["backend"]) == b"nccl"
    x = synth(41, length)
    x[length // 3 : length // 3 + 30000] = 0
    pa = W.PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
    ref = W.parakeet_log_mel(x, pa)[0]
    assert got["parakeet"].shape == ref.shape and np.abs(got["parakeet"] - ref).max() <= 5e-4  # SURVEY 8d tolerance
    refw = W.whisper_log_mel(x, 128)
    assert got["whisper"].shape == refw.shape and np.abs(got["whisper"] - refw).max() <= 1e-4
    assert refw.min() > refw.max() - 2.0 - 1e-6 and (refw == refw.min()).mean() > 0.02  # the clamp floor is reached (silence)
    clips = np.stack([synth(300 + i, 48000) * (0.2 + 0.1 * i) for i in range(7)])
    refc = np.stack([W.whisper_log_mel(c, 80) for c in clips])
    assert got["clips"].shape == refc.shape and np.abs(got["clips"] - refc).max() <= 1e-4
