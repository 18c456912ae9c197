import pytest

from baton_b200.control import (UpdateInProgress, UpdateManager, UpdateNotInProgress)
from conftest import run_async


@run_async
async def test_round_state_machine():
    um = UpdateManager("exp")
    assert um.update_name == "update_exp_00000"
    assert not um.in_progress and len(um) == 0
    with pytest.raises(UpdateNotInProgress):
        um.client_start("a")
    with pytest.raises(UpdateNotInProgress):
        um.client_end("a", {})
    await um.start_update(n_epoch=3)
    assert um.in_progress and um.update_meta == {"n_epoch": 3}
    with pytest.raises(UpdateInProgress):
        await um.start_update(n_epoch=1)
    um.client_start("a"), um.client_start("b")
    assert len(um) == 2 and um.clients_left == 2
    um.client_end("a", {"n_samples": 1})
    assert um.clients_left == 1
    resp = um.end_update()
    assert resp == {"a": {"n_samples": 1}}
    assert not um.in_progress and um.n_updates == 1 and len(um) == 0
    await um.start_update(n_epoch=1)
    assert um.update_name == "update_exp_00001"
    assert um.clients == set() and um.client_responses == {}
    um.end_update()
    assert len(um.round_times) == 2


@run_async
async def test_client_drop_and_snapshot_restore():
    um = UpdateManager("x")
    await um.start_update(n_epoch=1)
    um.client_start("a"), um.client_start("b")
    um.client_end("a", {})
    assert um.client_drop("b") is True and um.clients_left == 0
    assert um.client_drop("a") is False        # already responded: keep its update
    assert um.client_drop("zzz") is False
    um.end_update()
    um.loss_history.extend([1.0, 0.5])
    snap = um.snapshot()
    um2 = UpdateManager("x")
    um2.restore(snap)
    assert um2.n_updates == 1 and um2.loss_history == [1.0, 0.5]
    assert um2.update_name == "update_x_00001"
    st = um2.state()
    assert st["clients_left"] == 0 and st["in_progress"] is False


def test_default_name_is_random():
    assert UpdateManager().name != UpdateManager().name
    with pytest.raises(UpdateNotInProgress):
        UpdateManager().end_update()
