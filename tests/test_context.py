"""Rank layouts of every process-group kind, checked against the layouts documented in the reference docstrings
(``internlm/core/context/process_group_initializer.py``)."""
import pytest

from internevo_b200.core.context import Config, ParallelMode, ParallelSizes, group_rank_lists, layout_for_rank


def test_tensor_data_pipeline_layout():
    s = ParallelSizes(world=16, pipeline=2, tensor=2, zero1=2)
    assert group_rank_lists(ParallelMode.TENSOR, s)[:2] == [[0, 1], [2, 3]]
    assert group_rank_lists(ParallelMode.DATA, s) == [[0, 2, 4, 6], [1, 3, 5, 7], [8, 10, 12, 14], [9, 11, 13, 15]]
    assert group_rank_lists(ParallelMode.PIPELINE, s)[0] == [0, 8]
    assert group_rank_lists(ParallelMode.ZERO1, s) == [[0, 2], [1, 3], [4, 6], [5, 7], [8, 10], [9, 11], [12, 14], [13, 15]]


def test_zero1_full():
    s = ParallelSizes(world=16, pipeline=2, tensor=2, zero1=4)
    assert group_rank_lists(ParallelMode.ZERO1, s) == [[0, 2, 4, 6], [1, 3, 5, 7], [8, 10, 12, 14], [9, 11, 13, 15]]


def test_expert_layout():
    s = ParallelSizes(world=8, tensor=2, num_experts=2)
    assert group_rank_lists(ParallelMode.EXPERT, s) == [[0, 2], [4, 6], [1, 3], [5, 7]]
    assert group_rank_lists(ParallelMode.EXPERT_DATA, s) == [[0, 4], [2, 6], [1, 5], [3, 7]]


def test_isp_weight_layout():
    s = ParallelSizes(world=16, pipeline=2, tensor=2, weight=4, zero1=2, isp=True)
    assert group_rank_lists(ParallelMode.WEIGHT, s)[:2] == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert group_rank_lists(ParallelMode.WEIGHT_DATA, s)[:2] == [[0, 4], [1, 5]]
    z = group_rank_lists(ParallelMode.ZERO1, s)
    assert [0, 4] in z and [1, 5] in z


@pytest.mark.parametrize("world,pp,tp", [(8, 1, 1), (8, 2, 2), (8, 1, 8), (16, 4, 2)])
def test_every_rank_in_exactly_one_group(world, pp, tp):
    s = ParallelSizes(world=world, pipeline=pp, tensor=tp)
    for mode in (ParallelMode.TENSOR, ParallelMode.DATA, ParallelMode.PIPELINE, ParallelMode.ZERO1, ParallelMode.WEIGHT,
                 ParallelMode.WEIGHT_DATA):
        seen = sorted(r for g in group_rank_lists(mode, s) for r in g)
        assert seen == list(range(world)), (mode, seen)
    lay = layout_for_rank(0, s)
    assert lay[ParallelMode.TENSOR] == list(range(tp))


def test_config_attr_dict(tmp_path):
    f = tmp_path / "cfg.py"
    f.write_text("import os\nA = 1\nmodel = dict(hidden_size=8, nested=dict(x=2))\n")
    c = Config.from_file(str(f))
    assert c.A == 1 and c.model.hidden_size == 8 and c.model.nested.x == 2 and "os" not in c
    c.model._add_item("y", dict(z=3))
    assert c.model.y.z == 3


def test_nettest_zero3_and_gqa_layouts():
    s = ParallelSizes(world=10, nettest=4)
    assert group_rank_lists(ParallelMode.NETTEST, s) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]   # tail group is shorter
    s = ParallelSizes(world=8, tensor=2, zero1=2, fsdp=True)
    z3 = group_rank_lists(ParallelMode.ZERO3_DP, s)
    assert sorted(r for g in z3 for r in g) == list(range(8)) and all(len(g) == 2 for g in z3)
    for g in z3:   # members sit at the same position of different zero1 (fsdp) groups
        zs = [next(i for i, z in enumerate(group_rank_lists(ParallelMode.ZERO1, s)) if r in z) for r in g]
        assert len(set(zs)) == len(zs)
    from internevo_b200.core.context import modes_to_build

    assert ParallelMode.ZERO3_DP in modes_to_build(s) and ParallelMode.GQA in modes_to_build(s, gqa=True)
    s = ParallelSizes(world=16, tensor=16, gqa_q_heads=32, gqa_kv_heads=8)
    gq = group_rank_lists(ParallelMode.GQA, s)
    assert gq[0] == [0, 1] and len(gq) == 8   # 16 TP ranks, 8 kv heads: two ranks share one kv head


def test_seed_manager_modes_are_independent_streams():
    import torch

    from internevo_b200.core.context.random import SeedManager

    sm = SeedManager()
    sm.add_seed(ParallelMode.DATA, 1024)
    sm.add_seed(ParallelMode.TENSOR, 2048)
    sm.set_mode(ParallelMode.DATA)
    a1 = torch.rand(3)
    sm.set_mode(ParallelMode.TENSOR)
    b1 = torch.rand(3)
    sm.set_mode(ParallelMode.DATA)   # resumes the DATA stream where it stopped
    a2 = torch.rand(3)
    torch.manual_seed(1024)
    ref = torch.rand(6)
    assert torch.equal(torch.cat([a1, a2]), ref) and not torch.equal(a1, b1)
