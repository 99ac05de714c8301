"""One process, several devices (hdb_multi_*): a batch cut into contiguous model ranges must give the
results of the single-device call bit for bit.  On a box with one GPU the same device is listed several times, which
still exercises the sharding, the per-device handles and the host threads; the NCCL gather of the reduced results needs
distinct devices and runs when there are at least two."""
import numpy as np
import pytest

from hd_b200 import _lib, frames
from hd_b200.model import flatten_model

pytestmark = pytest.mark.gpu


def _case(N=8, nM=70, nLC=20):
    fm = flatten_model(frames.uniform_frame(N))
    return fm, frames.uniform_frame_sections(N, 7, nM), frames.uniform_frame_loads(N, 7, nM, nLC)


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_sharded_batch_is_bit_identical_to_one_device(n_shards):
    fm, secs, loads = _case()
    one = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
    ref = one.static_batch(secs, loads)
    one.close()
    n_dev = _lib.device_count()
    devices = [d % n_dev for d in range(n_shards)]
    mt = _lib.MultiTopology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, devices)
    got = mt.static_batch(secs, loads)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    # the ranges are contiguous, cover the batch and differ by at most one model
    sizes = [mt.shard(70, d)[1] for d in range(n_shards)]
    firsts = [mt.shard(70, d)[0] for d in range(n_shards)]
    assert sum(sizes) == 70 and max(sizes) - min(sizes) <= 1 and firsts == list(np.cumsum([0] + sizes[:-1]))
    mt.close()


def test_one_bad_model_stays_in_its_shard():
    fm, secs, loads = _case(nM=10)
    secs[7, 3, 1] = -1.0
    n_dev = _lib.device_count()
    mt = _lib.MultiTopology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, [d % n_dev for d in range(4)])
    status, fail, disp, forces, react = mt.static_batch(secs, loads)
    assert list(np.flatnonzero(status)) == [7] and status[7] == _lib.STATUS_NONFINITE
    mt.close()


def test_resident_shards_and_nccl_summary():
    torch = pytest.importorskip("torch")
    n_dev = _lib.device_count()
    if n_dev < 2:
        pytest.skip("the NCCL gather needs two distinct devices")
    fm, secs, loads = _case(nM=37)
    nLC = loads.shape[1]
    secs[30, 0, 0] = np.nan
    devices = list(range(min(n_dev, 4)))
    mt = _lib.MultiTopology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, devices)
    counts, d_sec, d_load, d_disp, d_st, d_fc = [], [], [], [], [], []
    for d in devices:
        m0, nm = mt.shard(37, d)
        dev = torch.device("cuda", d)
        counts.append(nm)
        d_sec.append(torch.from_numpy(secs[m0:m0 + nm]).to(dev))
        d_load.append(torch.from_numpy(loads[m0:m0 + nm]).to(dev))
        d_disp.append(torch.empty((nm, nLC, fm.nP, 3), dtype=torch.float64, device=dev))
        d_st.append(torch.zeros((nm,), dtype=torch.int32, device=dev))
        d_fc.append(torch.zeros((nm,), dtype=torch.int32, device=dev))
    mt.static_batch_dev(counts, d_sec, nLC, d_load, d_disp, None, None, d_st, d_fc)
    mt.synchronize()
    summary = mt.gather_summary(counts, d_st, d_disp, nLC)
    one = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
    with np.errstate(all="ignore"):
        status, fail, disp, _, _ = one.static_batch(secs, loads, want=("disp",))
    one.close()
    for d in devices:
        m0, nm = mt.shard(37, d)
        ok = status[m0:m0 + nm] == 0
        assert summary[d, 0] == np.count_nonzero(~ok) and summary[d, 2] == nm
        assert summary[d, 1] == np.abs(disp[m0:m0 + nm][ok]).max()
        assert np.array_equal(d_disp[d].cpu().numpy()[ok], disp[m0:m0 + nm][ok])
    mt.close()
