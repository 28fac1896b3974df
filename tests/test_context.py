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


def test_initializer_classes_are_views_of_the_layout_table():
    """The reference's ``Initializer_*`` API (one class per group kind, ``init_dist_group()`` → tuple) on top of the layout
    table: no process group is created without an initialised backend, the rank sets are the table's."""
    from internevo_b200.core.context.process_groups import (Initializer_Data, Initializer_Expert_Data, Initializer_Pipeline,
                                                            Initializer_Tensor, Initializer_Zero1, ParallelMode)

    args = dict(world_size=16, weight_parallel_size=1, weight_data_parallel_size=8, sequence_parallel_size=2,
                data_parallel_size=4, pipeline_parallel_size=2, tensor_parallel_size=2, zero1_parallel_size=2,
                nettest_parallel_size=32, expert_parallel_size=1)
    for rank in range(16):
        lr, n, group, cpu, ranks, mode = Initializer_Tensor(rank, **args).init_dist_group()
        assert mode is ParallelMode.TENSOR and ranks == [rank // 2 * 2, rank // 2 * 2 + 1] and ranks[lr] == rank and n == 2
        assert group is None and cpu is None
        lr, n, _, _, ranks, _ = Initializer_Pipeline(rank, **args).init_dist_group()
        assert ranks == [rank % 8, rank % 8 + 8] and ranks[lr] == rank
        lr, n, _, _, ranks, _ = Initializer_Data(rank, **args).init_dist_group()
        assert n == 4 and ranks == [rank // 8 * 8 + rank % 2 + 2 * k for k in range(4)]
        lr, n, _, _, ranks, _ = Initializer_Zero1(rank, **args).init_dist_group()
        assert n == 2 and rank in ranks and all((r - rank) % 2 == 0 for r in ranks)
    margs = dict(args, pipeline_parallel_size=1, tensor_parallel_size=1, sequence_parallel_size=1, data_parallel_size=16,
                 zero1_parallel_size=16, weight_data_parallel_size=16, expert_parallel_size=4)
    ep, edp = Initializer_Expert_Data(6, **margs).init_dist_group()
    assert ep[4] == [4, 5, 6, 7] and edp[4] == [2, 6, 10, 14] and ep[5] is ParallelMode.EXPERT


def _node_locality(rank, world, kw):
    import os

    os.environ["B200_NODE_ID"] = f"node{rank // 2}"      # two "nodes" of two ranks each on this box
    from common import tiny_config

    from internevo_b200.core.context import ParallelMode, global_context as gpc
    from internevo_b200.initialize import initialize_distributed_env
    from internevo_b200.parallel import fused, moe_fused, sp_attention, symm

    initialize_distributed_env(config=tiny_config(tp=2), launcher="torch", seed=1)
    assert symm.node_ids() == ["node0", "node0", "node1", "node1"], symm.node_ids()
    tensor, data = gpc.get_group(ParallelMode.TENSOR), gpc.get_group(ParallelMode.DATA)
    intra = (symm.group_is_intra_node(tensor), symm.group_is_intra_node(data), symm.group_is_intra_node(None))
    # with symmetric memory "present" the back-ends still refuse the group that leaves the node (and never touch CUDA for it)
    symm.symm_available = lambda: True
    refused = (symm.peer_addressable(data), fused.enable_tp(data), fused.isp_backend(data),
               moe_fused.backend_for(data, 64, 128, 4), sp_attention.backend_for(data, 64, 4, 2, 16))
    tensor_ok = symm.peer_addressable(tensor)
    # groups made outside gpc (no table): the group itself is asked once
    symm._node_ids = None
    asked = (symm.group_is_intra_node(tensor), symm.group_is_intra_node(data))
    return intra, refused, tensor_ok, asked


def test_peer_memory_backends_only_serve_groups_inside_one_node():
    """Multi-node jobs: tensor groups (ranks {0,1}, {2,3}) stay inside a node and may use peer memory; the data-parallel groups
    ({0,2}, {1,3}) cross nodes, so every peer-memory back-end declines them and their collectives stay on NCCL."""
    from common import run_distributed

    for intra, refused, tensor_ok, asked in run_distributed(_node_locality, 4, {}):
        assert intra == (True, False, True) and asked == (True, False)
        assert refused == (False, None, None, None, None)
        assert tensor_ok is True
