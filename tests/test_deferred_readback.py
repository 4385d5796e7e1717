"""Host logic of the deferred num_rendered read-back (lara_b200/rasterizer.py: _Pending, LazyCount, ForwardState,
_drain_ready), driven on the CPU with stand-ins for the CUDA event and the pinned slot: counts are read once, an
overflow re-runs stage 2 exactly once with a larger capacity and warns (unless the caller says nothing consumed the
outputs yet), the high-water mark feeds the next optimistic capacity, resolved entries leave the pending list."""
import warnings

import pytest
import torch

from lara_b200 import rasterizer as R


class FakeEvent:
    def __init__(self, done=True):
        self.done, self.syncs = done, 0

    def query(self):
        return self.done

    def synchronize(self):
        self.syncs += 1
        self.done = True


@pytest.fixture
def clean_state(monkeypatch):
    monkeypatch.setattr(R, "_capacity_hwm", {})
    monkeypatch.setattr(R, "_pending", {})
    monkeypatch.setattr(R, "_readback_pool", {})
    monkeypatch.setattr(R, "_dev_index", lambda device: 0)
    return torch.device("cpu")


def make_pending(dev, counts, capacity, done=True):
    slot = torch.zeros(R._SLOT_WORDS, dtype=torch.int32)
    slot[:len(counts)] = torch.tensor(counts, dtype=torch.int32)
    ev = FakeEvent(done)
    reruns = []
    p = R._Pending(dev, (slot, ev), len(counts), capacity, reruns.append)
    R._pending.setdefault(0, []).append(p)
    return p, ev, reruns


def test_counts_are_read_once_and_the_slot_goes_back_to_the_pool(clean_state):
    p, ev, reruns = make_pending(clean_state, [100, 250, 30], capacity=1000, done=False)
    assert not p.ready()
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert p.resolve() == [100, 250, 30]
        assert p.resolve() == [100, 250, 30]          # second call: cached, no second wait
    assert ev.syncs == 1 and reruns == []
    assert R._capacity_hwm[0] == 250                   # per-view maximum feeds the next optimistic capacity
    assert R._pending[0] == [] and len(R._readback_pool[0]) == 1
    assert R.initial_capacity(10, clean_state) == max(int(250 * 1.25) + 1024, 80, 1 << 16)


def test_overflow_reruns_stage_two_once_and_warns(clean_state):
    p, ev, reruns = make_pending(clean_state, [5000, 70000], capacity=65536)
    with pytest.warns(RuntimeWarning, match="exceeded the optimistic capacity"):
        counts = p.resolve()
    assert counts == [5000, 70000] and reruns == [int(70000 * 1.25) + 1024]
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        p.resolve()
    assert len(reruns) == 1


def test_eager_check_inside_the_forward_does_not_warn(clean_state):
    p, ev, reruns = make_pending(clean_state, [200000], capacity=65536)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        p.resolve(stale_ok=True)                       # debug=True / SRF_SYNC_NUM_RENDERED=1: nothing consumed yet
    assert reruns == [int(200000 * 1.25) + 1024]


def test_counts_above_int31_are_read_as_unsigned(clean_state):
    p, ev, reruns = make_pending(clean_state, [-2147483648 + 5], capacity=1 << 40)
    assert p.resolve() == [2147483653]


def test_drain_resolves_only_what_has_landed(clean_state):
    a, ev_a, _ = make_pending(clean_state, [10], capacity=100, done=True)
    b, ev_b, _ = make_pending(clean_state, [20], capacity=100, done=False)
    R._drain_ready(clean_state)
    assert a.counts == [10] and b.counts is None and R._pending[0] == [b]
    assert ev_b.syncs == 0                              # never blocks
    ev_b.done = True
    R._drain_ready(clean_state)
    assert b.counts == [20] and R._pending[0] == []


def test_lazy_count_and_forward_state(clean_state):
    p, ev, _ = make_pending(clean_state, [42, 7], capacity=100, done=False)
    st = R.ForwardState(None, None, None, None, 100, nviews=2, pending=p)
    n0, n1 = R.LazyCount(p, 0), R.LazyCount(p, 1)
    assert ev.syncs == 0                                # building the int-like objects does not wait
    assert int(n1) == 7 and n0 == 42 and f"{n0}" == "42" and list(range(50))[n0] == 42
    assert st.num_rendered == [42, 7] and st.resolve() == [42, 7] and ev.syncs == 1
    single = R.ForwardState(None, None, None, None, 0, 0)
    assert single.num_rendered == 0 and single.resolve() == [0]
