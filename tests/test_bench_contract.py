"""bench.py's host-side pieces that need no GPU: the config object both arms print, the reference arm
(compiled reference search, oracle/_ref, driving a network callback from host threads in both feeding
modes) on a tiny CPU network, and the shape of its result object."""
import argparse

import pytest
import torch

import bench
from tests import oracles


def test_both_arms_print_the_same_config_object():
    a = argparse.Namespace(games=4096, nn_batch=256, parts=2, opening_plies=16)
    for world in (1, 2, 8):
        ours = bench.selfplay_config(a, world, "net")
        ref = bench.selfplay_config(a, world, "net")
        assert ours == ref and ours["games_per_gpu"] == 4096 // world
        assert ("configs[2]" in ours["workload"]) == (world == 1) and ("configs[3]" in ours["workload"]) == (world > 1)


def test_defaults_are_the_self_play_workload(monkeypatch):
    seen = {}
    monkeypatch.setattr(bench, "run_selfplay", lambda args: seen.setdefault("ours", args) and 0)
    monkeypatch.setattr(bench, "run_reference_selfplay", lambda args: seen.setdefault("ref", args) and 0)
    monkeypatch.setattr(bench.os, "dup2", lambda a, b: None)
    monkeypatch.setattr("sys.argv", ["bench.py"])
    bench.main()
    monkeypatch.setattr("sys.argv", ["bench.py", "--impl", "reference", "--gpus", "2"])
    bench.main()
    assert seen["ours"].workload == "selfplay" and seen["ours"].games == 4096 and seen["ours"].steps == 20
    assert seen["ours"].warmup >= 3 and seen["ref"].gpus == 2


@pytest.mark.timeout(600)
def test_reference_arm_on_a_tiny_cpu_network():
    """the --impl reference arm end to end, minus the GPU: 800-rollout moves in 80-rollout slices on the
    compiled reference search, one call per wave and batched through the collector"""
    if not oracles.have_ref(19):
        pytest.skip("oracle/_ref not built")
    from elf_b200.model import FusedActor, PolicyValueNet

    torch.manual_seed(0)
    fa = FusedActor(PolicyValueNet(19, num_block=1, dim=8).eval(), batchsize=256, dtype=torch.float32, cuda_graph=False)
    r = bench.ref_selfplay(fa, torch.device("cpu"), steps=2, warmup=1)
    assert r["kind"] == "reference" and r["unit"] == "moves/s" and r["value"] > 0 and r["cores"] >= 1
    assert set(r["modes"]) == {"one_call_per_wave", "batched"} and r["mode"] in r["modes"]
    for m in r["modes"].values():
        assert m["value"] > 0 and m["nn_positions_per_s"] > 0
    assert r["value"] == max(m["value"] for m in r["modes"].values())
