"""Checkpoint / resume of a scenario and the asynchronous run entry.

The reference's loop state (src/greb.f90:226-234) is Ts1,Ta1,To1,q1,cap_surf + the step counter + the
monthly accumulators and tsmn.  A run that is saved after Y years (or in the middle of a year),
restored into a FRESH handle and continued must equal the uninterrupted run bit for bit."""
import numpy as np
import pytest

import greb_b200
from greb_b200 import host
from test_gpu_parity import make_ensemble, product_physics

pytestmark = pytest.mark.gpu


def _members():
    ps = [product_physics(), product_physics(kappa=9.3e5, a_cloud=0.33), product_physics(da_ice=0.29, ce=1.8e-3)]
    return ps, [np.full(6, c, dtype=np.float32) for c in (680.0, 420.0, 900.0)]


@pytest.mark.parametrize("arith", ["exact", "fast"])
def test_resume_at_a_year_boundary_is_bit_identical(forcing, tmp_path, arith):
    ps, co2 = _members()
    a = make_ensemble(forcing, ps, co2)
    a.set_arithmetic(arith)
    a.spinup(1)
    a.reset_scenario()
    full, gm_full, gc_full = a.run(6)
    end_full = a.get_states()
    a.close()

    b = make_ensemble(forcing, ps, co2)
    b.set_arithmetic(arith)
    b.spinup(1)
    b.reset_scenario()
    first, gm1, _ = b.run(3)
    path = str(tmp_path / "scenario_year3.npz")
    host.save_checkpoint(path, b)
    b.close()

    c = make_ensemble(forcing, ps, co2)                  # fresh handle: no spin-up, no earlier years
    c.set_arithmetic(arith)
    assert host.load_checkpoint(path, c) == 3 * 730 + 1
    assert c.get_calendar() == 3 * 730 + 1
    rest, gm2, gc2 = c.run(3)
    assert np.array_equal(first, full[:, :3]) and np.array_equal(rest, full[:, 3:])
    assert np.array_equal(gm1, gm_full[:, :3]) and np.array_equal(gm2, gm_full[:, 3:])
    assert np.array_equal(gc2, gc_full[:, 3:])
    assert np.array_equal(c.get_states(), end_full)
    c.close()


def test_resume_in_the_middle_of_a_year(forcing, tmp_path):
    """mid-month, mid-year checkpoint: the accumulators carry partial sums (src/greb.f90:145-149)"""
    ps, co2 = _members()
    a = make_ensemble(forcing, ps, co2)
    a.spinup(1)
    a.reset_scenario()
    full, gm_full, _ = a.run(2)
    end_full = a.get_states()
    a.close()

    b = make_ensemble(forcing, ps, co2)
    b.spinup(1)
    b.reset_scenario()
    b.run(1, want_output=False)
    b.time_steps(731, 333)                               # stops on 15 June of year 2, first half-day
    acc = b.get_accumulators()
    assert np.abs(acc[:, 0]).max() > 0 and np.abs(acc[:, 5]).max() > 0     # Tmm and tsmn hold partial sums
    path = str(tmp_path / "mid.npz")
    host.save_checkpoint(path, b)
    jan_may = [b.get_monthly(m)[:5] for m in range(3)]
    b.close()

    c = make_ensemble(forcing, ps, co2)
    it = host.load_checkpoint(path, c)
    assert it == 731 + 333
    c.time_steps(it, 2 * 730 + 1 - it)                   # to the year boundary
    assert c.get_calendar() == 2 * 730 + 1
    assert np.array_equal(c.get_states(), end_full)
    for m in range(3):
        got = c.get_monthly(m)                           # months written by THIS launch: June .. December
        assert np.array_equal(got[:7], full[m, 1, 5:12]), m
        assert np.array_equal(jan_may[m], full[m, 1, :5]), m
    with pytest.raises(greb_b200.GrebError):
        c.time_steps(1, 731)                             # more than a year per launch
    c.close()


def test_run_async_chained_equals_run(forcing):
    """run_async x 3 + one wait == run(3): same records in the caller's buffers, same end state; the
    checked-before-launch arguments leave the calendar untouched when they are wrong."""
    import torch
    ps, co2 = _members()
    a = make_ensemble(forcing, ps, co2)
    a.spinup(1)
    a.reset_scenario()
    want, gm, _ = a.run(3)
    end = a.get_states()
    a.close()

    b = make_ensemble(forcing, ps, co2)
    b.spinup(1)
    b.reset_scenario()
    bufs = [torch.empty((3, 1, 12, 5, 48, 96), dtype=torch.float32).pin_memory() for _ in range(3)]
    for y in range(3):
        b.run_async(1, bufs[y].data_ptr())
    assert b.get_calendar() == 3 * 730 + 1               # implicit wait
    for y in range(3):
        assert np.array_equal(bufs[y].numpy()[:, 0], want[:, y]), y
    assert np.array_equal(b.get_states(), end)
    # run_async without records + fetch_monthly_async behind a state transfer (the bench's e2e pipeline)
    c = make_ensemble(forcing, ps, co2)
    c.spinup(1)
    c.reset_scenario()
    st = torch.empty((3, 5, 48, 96), dtype=torch.float32).pin_memory()
    for y in range(3):
        c.run_async(1)
        c.get_states_async(st.data_ptr())
        c.fetch_monthly_async(bufs[y].data_ptr())
        c.sync_compute()
    c.wait()
    for y in range(3):
        assert np.array_equal(bufs[y].numpy()[:, 0], want[:, y]), y
    assert np.array_equal(st.numpy(), end)
    c.close()
    it = b.get_calendar()
    with pytest.raises(greb_b200.GrebError):
        b.run(1, out_members=[0, 7])                     # bad member index: nothing may have been launched
    assert b.get_calendar() == it and np.array_equal(b.get_states(), end)
    out, gm2, _ = b.run(1, out_members=[2, 0])
    assert out.shape[0] == 2 and np.all(np.isfinite(gm2))
    b.close()


def test_failed_init_leaves_the_handle_unusable_not_dangling(forcing):
    """a second init that fails must not leave inited = true with freed device pointers"""
    e = make_ensemble(forcing, [product_physics()], [[680.0]])
    e.spinup(1)
    e.set_member(0, product_physics(kappa=9e6), [680.0])  # needs more helper rows than the kernel has
    with pytest.raises(greb_b200.GrebError):
        e.init()
    with pytest.raises(greb_b200.GrebError):
        e.spinup(1)
    with pytest.raises(greb_b200.GrebError):
        e.run(1)
    e.set_member(0, product_physics(), [680.0])
    e.init()
    e.spinup(1)
    e.close()


def test_ensemble_field_moments_on_the_device(forcing):
    """greb_b200_ensemble_moments: sum and sum of squares of the year's records over the members, float64"""
    from greb_b200 import sharding
    ps, co2 = _members()
    e = make_ensemble(forcing, ps + ps[:2], co2 + co2[:2])          # 5 members: exercises the unrolled loop + tail
    e.spinup(1)
    e.reset_scenario()
    out, _, _ = e.run(1)
    s, q = e.ensemble_moments()
    o = out[:, 0].astype(np.float64)
    assert np.allclose(s, o.sum(0), rtol=1e-15) and np.allclose(q, (o * o).sum(0), rtol=1e-14)
    mean, var, cnt = sharding.reduce_field_moments(e, e.n)         # single rank: no collective
    assert cnt == 5 and np.allclose(mean, o.mean(0), rtol=1e-14) and np.allclose(var, o.var(0), rtol=1e-6, atol=1e-9)
    e.close()


def test_async_state_upload_and_device_views(forcing):
    """greb_b200_set_states_async (pinned upload behind the compute stream), greb_b200_diag_device and
    greb_b200_ensemble_moments_device (the vectors a multi-GPU driver reduces) against their host-side twins"""
    import torch
    ps, co2 = _members()
    a = make_ensemble(forcing, ps, co2)
    a.spinup(1)
    a.reset_scenario()
    _, gm1, gmc1 = a.run(1)
    st1 = a.get_states()
    out2, gm2, gmc2 = a.run(1)
    # rewind: upload the year-1 states through the asynchronous entry, set the calendar, run year 2 again
    pin = torch.from_numpy(st1.copy()).pin_memory()
    a.set_states_async(pin.data_ptr())
    a.wait()
    assert np.array_equal(a.get_states(), st1)
    a.set_calendar(730 + 1)
    out2b, gm2b, gmc2b = a.run(1)
    assert np.array_equal(out2b, out2) and np.array_equal(gm2b, gm2) and np.array_equal(gmc2b, gmc2)

    class _View:
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

    ptr, n = a.diag_device()
    assert n == 2 * a.n
    d = torch.as_tensor(_View(ptr, n, "<f4"), device="cuda:0").cpu().numpy().reshape(a.n, 2)
    assert np.array_equal(d[:, 0], gm2b[:, -1]) and np.array_equal(d[:, 1], gmc2b[:, -1])
    s, q = a.ensemble_moments()
    pS, pQ, ne = a.ensemble_moments_device()
    assert ne == 12 * 5 * 48 * 96
    assert np.array_equal(torch.as_tensor(_View(pS, ne, "<f8"), device="cuda:0").cpu().numpy(), s.ravel())
    assert np.array_equal(torch.as_tensor(_View(pQ, ne, "<f8"), device="cuda:0").cpu().numpy(), q.ravel())
    assert not a.flags().any()
    a.close()
