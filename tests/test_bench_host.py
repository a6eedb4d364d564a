"""Host-side helpers of bench.py that can be checked without a GPU."""
import os
import sys
import types

import bench


def _fake_nvml(words):
    m = types.ModuleType("pynvml")
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByUUID = lambda u: (_ for _ in ()).throw(RuntimeError("no uuid"))
    m.nvmlDeviceGetHandleByIndex = lambda i: i
    m.nvmlDeviceGetCpuAffinity = lambda h, n: list(words)[:n] + [0] * max(0, n - len(words))
    return m


def test_bind_near_gpu_restricts_and_restores(monkeypatch):
    cur = os.sched_getaffinity(0)
    if len(cur) < 2:
        import pytest
        pytest.skip("one CPU only")
    keep = sorted(cur)[: len(cur) // 2]
    word = 0
    for c in keep:
        assert c < 64
        word |= 1 << c
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml([word]))
    try:
        old, note = bench.bind_near_gpu(0)
        assert old == cur and os.sched_getaffinity(0) == set(keep) and "local to the GPU" in note
    finally:
        os.sched_setaffinity(0, cur)
    assert os.sched_getaffinity(0) == cur


def test_bind_near_gpu_never_raises_and_leaves_the_mask_alone(monkeypatch):
    cur = os.sched_getaffinity(0)
    # a CPU set that does not intersect ours, the set we already have, and no NVML at all
    far = 0
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml([far]))
    assert bench.bind_near_gpu(0)[0] is None and os.sched_getaffinity(0) == cur
    word = 0
    for c in cur:
        word |= 1 << c
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml([word & (2 ** 64 - 1)]))
    assert bench.bind_near_gpu(0)[0] is None and os.sched_getaffinity(0) == cur
    broken = types.ModuleType("pynvml")
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    old, note = bench.bind_near_gpu(0)
    assert old is None and note.startswith("unbound") and os.sched_getaffinity(0) == cur
