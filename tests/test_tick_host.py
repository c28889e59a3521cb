"""The pair instantiation of the tick (two legs per packed FP32 instruction, csrc/qs_packed.cuh) against the plain scalar
templates, both on the HOST: tools/tick_host_check.cu is built twice (default and -DQS_PACK_LEGS=0) and the two must agree
to fp64 rounding over airborne, touching and standing sequences (contacts, warm starts, friction, early exits).  This is
what says that the packed kernels run the same arithmetic as the templates the fp64 GPU check and the flop model use."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _build_and_run(tmp, tag, flags):
    exe = os.path.join(tmp, f"tick_host_check_{tag}")
    res = subprocess.run([NVCC, "-std=c++17", "-w", *flags, "-o", exe, os.path.join(ROOT, "tools", "tick_host_check.cu")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    o64, o32 = os.path.join(tmp, f"{tag}_64.txt"), os.path.join(tmp, f"{tag}_32.txt")
    res = subprocess.run([exe, o64, o32], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    return o64, o32


def _rows(path):
    return [line.split() for line in open(path)]


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_pair_instantiation_is_the_scalar_arithmetic(tmp_path):
    tmp = str(tmp_path)
    p64, p32 = _build_and_run(tmp, "pairs", [])
    s64, s32 = _build_and_run(tmp, "scalar", ["-DQS_PACK_LEGS=0"])
    a, b = _rows(p64), _rows(s64)
    assert len(a) == len(b) and len(a) > 500
    worst = 0.0
    contact_rows = 0
    for x, y in zip(a, b):
        assert x[:3] == y[:3], "tick result / contact mask / invalid count differ"
        assert x[-2:] == y[-2:], "contact and PGS sweep counters differ: the early exit took another turn"
        worst = max(worst, float(np.abs(np.array(x[3:-2], float) - np.array(y[3:-2], float)).max()))
        contact_rows += int(x[1]) == 15
    assert contact_rows >= 100, "the standing sequences never got four feet down"
    assert worst < 1e-10, worst          # fp64: the same arithmetic up to the order of a few sums
    # fp32: the same code in single precision stays close over 400 ticks of contact (rounding differs: packed sums are
    # taken in another order), the flags of the first ticks agree
    a, b = _rows(p32), _rows(s32)
    assert len(a) == len(b)
    d = max(float(np.abs(np.array(x[3:-2], float) - np.array(y[3:-2], float)).max()) for x, y in zip(a, b))
    assert d < 5e-3, d
