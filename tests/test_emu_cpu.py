"""Kernel-logic tests without a GPU: the product's CUDA sources compiled with g++ against tests/cpu_emu/shim (blocks
run sequentially, threads are fibers, warp primitives and barriers are rendezvous) and run through the same C-ABI and
the same parity assertions as tests/test_gpu_parity.py.

This is TEST INFRASTRUCTURE: it checks indexing, list formats, reductions and the parity arithmetic of the kernels in
the build container, where no GPU exists.  It says nothing about performance, and the emulated library is never
loaded by the product (ddcmd_b200.lib() only ever opens libddcmd_b200.so, which needs a device).
"""
import ctypes
import os
import sys

import numpy as np
import pytest

import ddcmd_b200 as dd
import test_gpu_parity as tg
import test_zzzzzzz_variants as tv

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_emu"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu_handle():
    return dd._declare(ctypes.CDLL(build_emu.build()))


@pytest.fixture()
def emu(emu_handle, monkeypatch):
    dd.lib()                       # make sure the real library is what gets restored afterwards
    monkeypatch.setattr(dd, "_lib", emu_handle)
    return emu_handle


SMALL = ["popc_small", "ras_small", "tiny2"]


@pytest.mark.parametrize("name", tg.DECKS)
def test_emu_cells_bit_exact(emu, golden_dir, name):
    tg.test_cells_bit_exact(golden_dir, name)


@pytest.mark.parametrize("name", tg.DECKS)
def test_emu_pair_membership_bit_exact(emu, golden_dir, name):
    tg.test_pair_membership_bit_exact(golden_dir, name)


@pytest.mark.parametrize("name", tg.DECKS)
def test_emu_step0_forces_energy_virial(emu, golden_dir, name):
    tg.test_step0_forces_energy_virial(golden_dir, name)


@pytest.mark.parametrize("name", ["popc_small"])
def test_emu_per_bead_walk_bound(emu, golden_dir, name, monkeypatch):
    tv.test_per_bead_walk_bound_is_bitwise_neutral(golden_dir, name, monkeypatch)


@pytest.mark.parametrize("name", ["popc_small"])
def test_emu_pruned_rows(emu, golden_dir, name, monkeypatch):
    tv.test_pruned_rows_are_bitwise_neutral(golden_dir, name, monkeypatch)


@pytest.mark.parametrize("mode", ["barostat", "ur0"])
def test_emu_pruned_rows_other_modes(emu, golden_dir, mode, tmp_path, monkeypatch):
    tv.test_pruned_rows_with_barostat_and_displacement_rebuilds(golden_dir, mode, tmp_path, monkeypatch)


def test_emu_row_capacity_regrow(emu, golden_dir, monkeypatch):
    tv.test_row_capacity_regrow(golden_dir, monkeypatch)


def test_emu_near_edge_knob(emu, golden_dir, monkeypatch):
    tv.test_near_edge_knob_keeps_the_pair_set(golden_dir, monkeypatch)


@pytest.mark.parametrize("name", SMALL)
def test_emu_trajectory_40_steps(emu, golden_dir, name):
    tg.test_trajectory_40_steps(golden_dir, name)


def test_emu_printinfo_line(emu, golden_dir):
    tg.test_printinfo_line_matches_reference_data_file(golden_dir)


def test_emu_displacement_triggered_rebuild(emu, golden_dir, tmp_path):
    """DDC updateRate = 0: rebuild loops, pair counts and energies against the reference's trace (k_nbr_rbar_partial, k_nbr_check)."""
    import test_zzzz_ur0 as tu
    tu.check_ur0(golden_dir, "popc_small", tmp_path, 40)


@pytest.mark.parametrize("deck,variant", [("popc_small", "full"), ("ras_small", "full"), ("popc_small", "lang")])
def test_emu_nglfconstraint(emu, golden_dir, tmp_path, deck, variant):
    """NGLFCONSTRAINT: Langevin groups (bit-equal LCG64 streams), velocity constraints, barostat box trace (k_nglfc, k_constraint)."""
    import test_zzzzz_nglfc as tn
    tn.check_nglfc(golden_dir, deck, variant, tmp_path, 40 if deck == "popc_small" else 22)


def test_emu_nglf_with_langevin_groups(emu, golden_dir, tmp_path):
    import test_zzzzz_nglfc as tn
    tn.test_nglf_with_langevin_groups_is_the_unconstrained_pass(golden_dir, tmp_path)


@pytest.mark.parametrize("deck,variant", [("popc_small", None), ("popc_small", "full")])
def test_emu_simulateMaster(emu, golden_dir, tmp_path, deck, variant):
    """ddcb200_simulateMaster: data file, checkpoint cadence, ddcMD_CMDS and restart continuation against the reference's run."""
    import test_zzzzzz_master as tm
    tm.check_master(golden_dir, deck, variant, tmp_path)


def test_emu_subsetWrite(emu, golden_dir, tmp_path):
    import test_zzzzzz_master as tm
    tm.check_subset(golden_dir, tmp_path)


def test_emu_kinetic_classes(emu, golden_dir, tmp_path):
    import test_zzzzzz_master as tm
    tm.check_kinetic_classes(golden_dir, tmp_path)


def test_emu_paircorrelation(emu, golden_dir, tmp_path):
    import test_zzzzzz_master as tm
    tm.check_paircorr(golden_dir, tmp_path)


def test_emu_fullsize_checks_on_a_small_generated_deck(emu, tmp_path, monkeypatch):
    """The live-oracle comparison and the size-independent property checks of tests/test_zzz_fullsize.py, on the synthetic
    generator's ~4k-bead configuration so the emulation finishes in seconds."""
    import test_zzz_fullsize as tf
    monkeypatch.setattr(tf, "CACHE", str(tmp_path))
    if os.path.exists(tf.REF_DUMP):
        assert tf.check_against_live_oracle("popc_small") > 3000
        assert tf.check_against_live_oracle("popc_small", hashed=True) > 3000     # the million-bead form of the pair-list check
    assert tf.check_properties("popc_small") > 3000


def _torchrun(nproc, port, script, *args, env=None):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", script)] + list(args)
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e)


@pytest.mark.parametrize("name", ["popc_small", "tiny2"])
def test_emu_two_ranks_match_reference(emu_handle, name):
    """The ddc decomposition (re-domain, halo lists, ghost halo per step, all-reduced energyInfo) on 2 emulated ranks,
    NCCL replaced by a shared-memory stand-in, against the single-rank reference outputs.  tiny2: bricks barely wider than
    the list radius, so nearly every bead is somebody's ghost."""
    r = _torchrun(2, 29561 if name == "popc_small" else 29563, "mgpu_worker.py", name, env={"DDCB200_TEST_EMU": "1"})
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("nproc,name,lattice", [(4, "ras_small", (4, 1, 1)), (8, "popc_small", (4, 2, 1))])
def test_emu_four_and_eight_ranks_match_reference(emu_handle, nproc, name, lattice):
    """The same worker on 4 and 8 emulated ranks: emigrants to non-adjacent bricks, rank pairs that exchange nothing (4x2x1 is
    the lattice of the 8-GPU scaling run), molecules with bonded terms and restraints crossing brick faces."""
    r = _torchrun(nproc, 29564 + nproc, "mgpu_worker.py", name, *[str(x) for x in lattice], env={"DDCB200_TEST_EMU": "1"})
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_emu_two_ranks_nglfconstraint(emu_handle):
    """NGLFCONSTRAINT (LANGEVIN groups, velocity constraints, barostat) on 2 emulated ranks against the single-rank reference trace:
    the per-bead random streams follow migrating beads bit for bit, the barostat works on the all-reduced virial."""
    r = _torchrun(2, 29571, "mgpu_nglfc_worker.py", "ras_small", "full", env={"DDCB200_TEST_EMU": "1", "DDCB200_NGLFC_STEPS": "22"})
    assert r.returncode == 0 and "MGPU_NGLFC_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_emu_bench_contract_two_ranks(emu_handle):
    """bench.py's multi-rank arm end to end (rank plumbing, max-over-ranks timing, one JSON line from rank 0)."""
    import json
    r = _torchrun(2, 29562, "emu_bench_check.py", "--gpus", "2", "--workload", "popc_small", "--steps", "3", "--warmup", "3",
                  "--equil-rounds", "1", "--equil-steps", "2")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "run_info", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in j, key
    assert j["n_gpus"] == 2 and j["steps"] == 3 and j["gpu_launches"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(j["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(j["e2e"])


def test_emu_is_not_the_product():
    """The product wrapper opens only libddcmd_b200.so; nothing under ddcmd_b200/ names the emulation."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "ddcmd_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "cpu_emu" not in text or f in ("engine.cuh",), f
                assert "libddcmd_b200_emu" not in text, f
