"""GPU parity of the DDCB200_GROUP pair-path variants (merged group rows, k_pair_group<2|4>): same assertions as
tests/test_gpu_parity.py.  Kept in its own file, collected last, so the default path's verdict never depends on it."""
import pytest

import test_gpu_parity as tg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", [2, 4])
@pytest.mark.parametrize("name", tg.DECKS)
def test_group_rows_membership_forces_energy(golden_dir, name, group, monkeypatch):
    monkeypatch.setenv("DDCB200_GROUP", str(group))
    tg.test_pair_membership_bit_exact(golden_dir, name)
    tg.test_step0_forces_energy_virial(golden_dir, name)


@pytest.mark.parametrize("group", [2, 4])
@pytest.mark.parametrize("name", tg.DECKS)
def test_group_rows_trajectory(golden_dir, name, group, monkeypatch):
    monkeypatch.setenv("DDCB200_GROUP", str(group))
    tg.test_trajectory_40_steps(golden_dir, name)
