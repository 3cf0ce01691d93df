"""CPU-only checks of bench.py's host logic: both arms describe the same workload with the same config dict (the driver
compares them), the whole-step algorithmic-byte formula is SURVEY section 8d's, and the 3-camera rig is three views."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import helpers as Hh  # noqa: E402,F401  (path set-up)


def test_config_is_identical_in_both_arms_and_names_the_workload():
    wl = bench.WORKLOADS["kitti-375x1242-1M"]
    a = bench.make_config("kitti-375x1242-1M", wl, 1, 1)
    b = bench.make_config("kitti-375x1242-1M", wl, 1, 1)
    assert a == b and a["workload"] == "kitti-375x1242-1M" and a["image"] == [375, 1242] and a["gaussians"] == 1_000_000
    assert "model" not in a
    assert bench.make_config("kitti-375x1242-1M", wl, 8, 1)["views_per_step"] == 8
    rig = bench.WORKLOADS["waymo-3cam-1066x1600-3M"]
    assert bench.make_config("waymo-3cam-1066x1600-3M", rig, 1, 3)["views_per_step"] == 3


def test_algorithmic_bytes_is_the_survey_formula():
    # B_alg = N_s*1220 + N_o*2790 + R*404 + Px*84 per view; the verdict's recomputation for configs[1]
    total, stages = bench.algorithmic_bytes(750_000, 250_000, 2_413_910, 375 * 1242)
    assert total == 750_000 * 1220 + 250_000 * 2790 + 2_413_910 * 404 + 465_750 * 84 == 2_626_842_640
    assert stages["blend_backward"] == 2_413_910 * 172 + 465_750 * 44


def test_rig_workload_is_three_cameras_sharing_one_timestep():
    wl = bench.WORKLOADS["waymo-3cam-1066x1600-3M"]
    views = bench.views_of_step(wl, 0, "cpu")
    assert len(views) == 3
    times = {t for _, t, _ in views}
    assert len(times) == 1
    # the three optical axes are 45 degrees apart (scene/dataset_readers.py:261-357: front-left / front / front-right)
    axes = [np.asarray(cam.world_view_transform)[:3, 2] for cam, _, _ in views]
    for a, b in ((axes[0], axes[1]), (axes[1], axes[2])):
        ang = math.degrees(math.acos(float(np.clip(np.dot(a, b), -1, 1))))
        assert abs(ang - 45.0) < 1e-3
    assert len(bench.views_of_step(bench.WORKLOADS["kitti-375x1242-1M"], 3, "cpu")) == 1


def test_elementwise_metric_catches_what_the_max_norm_hides():
    import torch
    b = torch.tensor([1.0, 1e-3, 1e-3])
    a = torch.tensor([1.0, 1e-3, 2e-3])             # a 100 % error on a small entry
    assert abs(Hh.rel_err(a, b) - 1e-3) < 1e-9      # "1e-3 relative" in the max norm ...
    frac, worst = Hh.elementwise_err(a, b, rtol=1e-4, atol_frac=1e-6)
    assert abs(frac - 1 / 3) < 1e-12 and worst > 500   # ... but one element in three is off by ~900x its bound
    assert Hh.elementwise_err(b, b)[0] == 0.0


class _FakeCuda:
    """Stand-in for torch.cuda that records what bench.InputStager asks of it (tensors live on the CPU)."""

    def __init__(self):
        self.log = []
        self.current = "compute"
        outer = self

        class Event:
            def __init__(self):
                self.recorded_on = None

            def record(self, stream=None):
                self.recorded_on = getattr(stream, "name", stream)
                outer.log.append(("record", self.recorded_on))

        class StreamCtx:
            def __init__(self, stream):
                self.stream = stream

            def __enter__(self):
                self.prev, outer.current = outer.current, self.stream.name

            def __exit__(self, *exc):
                outer.current = self.prev
                return False

        self.Event = Event
        self._ctx = StreamCtx

    def stream(self, s):
        return self._ctx(s)

    def current_stream(self, device=None):
        return type("S", (), {"name": self.current})()


class _FakeStream:
    def __init__(self, name, cuda):
        self.name, self.cuda = name, cuda

    def wait_event(self, ev):
        self.cuda.log.append(("wait", self.name, ev))


def test_input_stager_rotates_preallocated_slots_and_orders_reuse_behind_the_reader():
    import torch
    cuda = _FakeCuda()
    copy_stream = _FakeStream("copy", cuda)
    cam, cot = torch.arange(35.0), torch.arange(1000.0)
    st = bench.InputStager([cam, cot], "cpu", copy_stream, ring=True, slots=4, cuda=cuda)
    assert len(st.buffers) == 4 and all(b[0].shape == cam.shape and b[1].shape == cot.shape for b in st.buffers)
    ptrs = [(b[0].data_ptr(), b[1].data_ptr()) for b in st.buffers]
    seen, releases = [], []
    for step in range(9):
        cam.fill_(float(step))           # this step's host data
        cot.fill_(float(-step))
        cuda.log.clear()
        k, (dcam, dcot), (e_cam, e_cot) = st.stage()
        assert k == step % 4 and (dcam.data_ptr(), dcot.data_ptr()) == ptrs[k]      # no allocation after __init__
        assert torch.equal(dcam, cam) and torch.equal(dcot, cot)
        assert e_cam.recorded_on == "copy" and e_cot.recorded_on == "copy"
        waits = [x for x in cuda.log if x[0] == "wait"]
        if step < 4:
            assert waits == []                               # fresh slot: nothing to wait for
        else:
            # the copy stream waits for the event recorded on the COMPUTE stream when this slot's last reader was
            # queued, before it overwrites the slot
            assert len(waits) == 1 and waits[0][1] == "copy" and waits[0][2] is releases[step - 4]
            assert cuda.log.index(waits[0]) == 0
        assert cuda.current == "compute"                     # the stream context was left
        st.release(k)
        releases.append(st.slot_free[k])
        assert releases[-1].recorded_on == "compute"
        seen.append(k)
    assert seen == [0, 1, 2, 3, 0, 1, 2, 3, 0]
    # skipped input (diagnosis mode): no copy, but its event still exists
    k, (dcam, dcot), evs = st.stage(skip=(1,))
    assert dcot is None and dcam is not None and len(evs) == 2


def test_input_stager_alloc_mode_makes_a_tensor_per_step_and_marks_its_reader():
    import torch
    cuda = _FakeCuda()
    st = bench.InputStager([torch.ones(8)], "cpu", _FakeStream("copy", cuda), ring=False, cuda=cuda)
    assert st.buffers == []
    k, (d,), (ev,) = st.stage()
    assert torch.equal(d, torch.ones(8)) and ev.recorded_on == "copy"
    marked = []
    d_proxy = type("T", (), {"record_stream": lambda self, s: marked.append(s.name)})()
    st.reads_on_current_stream(d_proxy)
    assert marked == ["compute"]
    st.release(k)
    assert st.slot_free[k] is None                       # nothing to order in this mode
