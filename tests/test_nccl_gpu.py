"""World-size-2 NCCL run of the sharded scoring path: doc-range shards on two GPUs, one
ncclAllGather of the per-shard top-k + topk_merge_kernel, against the CPU oracle on the full corpus
(the gloo test checks numpy merge_topk, this one the device path).  Skips with < 2 GPUs."""
import ctypes
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def gpu_count():
    from searcharray_b200 import _lib
    n = ctypes.c_int(0)
    _lib.check(_lib.lib().sa_device_count(ctypes.byref(n)))
    return n.value


def test_two_gpu_allgather_topk_matches_oracle():
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    key = f"/tmp/sa_b200_test_uid_{os.getpid()}.bin"
    if os.path.exists(key):
        os.remove(key)
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_nccl_worker.py"), str(r), "2", key],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=900)
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        if os.path.exists(key):
            os.remove(key)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    assert "NCCL_OK 2" in outs[0], outs[0][-2000:]
