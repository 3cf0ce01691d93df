"""Host logic of the sync-free binning path (adgs_b200/gaussian_model.py:_resolve_counter_checks): which counter
read-backs a forward waits for, how the arena bound follows num_rendered, and that an overflow is reported. Fake events,
no GPU."""
import warnings

import torch

from adgs_b200 import scenes
from adgs_b200.gaussian_model import GaussianModel


class FakeEvent:
    def __init__(self, done):
        self.done = done
        self.waited = False

    def query(self):
        return self.done

    def synchronize(self):
        self.waited = True
        self.done = True


def _model():
    return GaussianModel(3, scenes.BENCH_ORDER_ARGS)


def _counters(num_rendered, overflow=0):
    return torch.tensor([num_rendered, overflow], dtype=torch.int32)


def test_blocking_resolve_waits_for_every_check_and_grows_the_arena():
    m = _model()
    evs = [FakeEvent(False), FakeEvent(True)]
    m._defer_counter_check(_counters(1000), evs[0], 10_000_000)
    m._defer_counter_check(_counters(5_000_000), evs[1], 10_000_000)
    m._resolve_counter_checks(block=True)
    assert evs[0].waited and not evs[1].waited          # only the unfinished one costs a wait
    assert m.__dict__["_counter_checks"] == []
    assert m._binning_capacity == int(1.3 * 5_000_000) + 65536
    assert m._last_num_rendered == 5_000_000


def test_non_blocking_resolve_only_takes_what_has_landed():
    m = _model()
    evs = [FakeEvent(True), FakeEvent(False)]
    m._defer_counter_check(_counters(2000), evs[0], 1 << 20)
    m._defer_counter_check(_counters(900_000), evs[1], 1 << 20)
    m._resolve_counter_checks(block=False)
    assert not evs[1].waited and len(m.__dict__["_counter_checks"]) == 1
    assert m._binning_capacity == int(1.3 * 2000) + 65536
    evs[1].done = True
    m._resolve_counter_checks(block=False)
    assert m.__dict__["_counter_checks"] == [] and m._binning_capacity == int(1.3 * 900_000) + 65536


def test_outstanding_forwards_are_not_waited_for():
    m = _model()
    evs = [FakeEvent(False), FakeEvent(False), FakeEvent(False)]
    for i, e in enumerate(evs):
        m._defer_counter_check(_counters(1000 * (i + 1)), e, 1 << 20)
    m._resolve_counter_checks(block=True, outstanding=1)
    assert evs[0].waited and evs[1].waited and not evs[2].waited
    assert len(m.__dict__["_counter_checks"]) == 1
    m._resolve_counter_checks(block=True, outstanding=1)     # the newest one stays in flight
    assert not evs[2].waited
    m._resolve_counter_checks(block=True)
    assert evs[2].waited and m.__dict__["_counter_checks"] == []
    assert m.sync_free_outstanding == 0                       # the shipped default waits for all of them


def test_overflow_is_reported_and_the_arena_enlarged():
    m = _model()
    m._binning_capacity = 64
    m._defer_counter_check(_counters(250_000, 1), FakeEvent(True), 64)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._resolve_counter_checks(block=True)
    assert any("overflow" in str(x.message) for x in w)
    assert m._binning_capacity == int(1.3 * 250_000) + 65536
    # a num_rendered above the capacity the forward ran with counts as an overflow even without the device flag
    m._defer_counter_check(_counters(2_000_000, 0), FakeEvent(True), 1_000_000)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._resolve_counter_checks(block=False)
    assert any("overflow" in str(x.message) for x in w)
