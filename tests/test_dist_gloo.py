"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: chain building, LPT partition, the single all-gather of
the scale vectors, replica replay, and the exact-mode sweep count.  The arithmetic is executed by tests/fakelib.py (the
oracle); the -m gpu suite covers the kernels, and bench.py --gpus N the NCCL path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _nw(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import torch.nn as nn
    import fakelib
    from dfq_b200 import workload, dist as ddist
    from dfq_b200.utils import layer_transform as LT
    from dfq_b200.utils.relation import create_relation
    torch.set_num_threads(2)
    fakelib.install_plain(fakelib.torch_sqrt)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        topo = workload.load_topology(os.path.join(GOLD, "topology_mobilenetv2.json"))
        graph, bottoms, _ = workload.build_graph(topo, seed=0)
        targ = [nn.Conv2d, nn.Linear]
        LT.merge_batchnorm(None, graph, bottoms, targ)
        rels = create_relation(graph, bottoms, targ)
        info = ddist.sharded_cross_layer_equalization(graph, rels, targ, mode=mode)
        keys = list(graph.keys())
        out = {"sweeps": np.array(info["sweeps"]), "owner": np.array(info["owner"])}
        for i, k in enumerate(keys):
            if type(graph[k]) in targ:
                out["w%d" % i] = graph[k].weight.detach().numpy()
                if graph[k].bias is not None:
                    out["b%d" % i] = graph[k].bias.detach().numpy()
        for i, r in enumerate(rels):
            out["S%d" % i] = r.S.numpy()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["exact", "per_chain"])
def test_two_rank_sharded_equalization(mode, tmp_path):
    port = 29500 + (os.getpid() % 2000) + (7 if mode == "exact" else 0)
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    gold = np.load(os.path.join(GOLD, "ref_mobilenetv2.npz"))
    # both ranks end with the same full model
    assert set(r0.files) == set(r1.files)
    assert set(r0["owner"].tolist()) == {0, 1}, "chains must be spread over both ranks"
    for k in r0.files:
        if k[0] in "wbS":
            assert _nw(r0[k], r1[k]) < 1e-5, k
    # ... and it is the reference's result
    n_rel = gold["relations"].shape[0]
    for i in range(n_rel):
        assert _nw(r0["S%d" % i], gold["S_%d" % i]) < 1e-5, i
    targets = gold["targets"]
    for j, pos in enumerate(targets):
        w = r0["w%d" % pos]
        assert abs(np.abs(w).max() - gold["cle_w_absmax"][j]) <= 1e-5 * gold["cle_w_absmax"][j]
        if "cle_bias_%d" % pos in gold.files and "b%d" % pos in r0.files:
            assert _nw(r0["b%d" % pos], gold["cle_bias_%d" % pos]) < 1e-5
    if mode == "exact":
        assert int(r0["sweeps"]) == int(gold["n_sweeps"]) == int(r1["sweeps"])


def test_partition_and_chains_are_deterministic():
    from dfq_b200.dist import build_chains, partition_lpt, _replay_exit_rule
    from dfq_b200.utils.relation import Relation
    rels = [Relation("a", "b", "x"), Relation("b", "c", "y"), Relation("d", "e", "z"), Relation("c", "f", "w")]
    assert build_chains(rels) == [[0, 1, 3], [2]]
    assert partition_lpt([5, 3, 3, 2], 2) == [0, 1, 1, 0]
    assert partition_lpt([1, 1, 1], 1) == [0, 0, 0]
    # dfq.py:81-115 replay: stops when diff <= thres or after converge_count stagnant sweeps
    assert _replay_exit_rule([1.0, 0.5, 1e-8], 2e-7, 20) == 3
    assert _replay_exit_rule([1.0] * 30, 2e-7, 3) == 4
    # a recording that ends before the rule fires says so (None) and can be continued by the next recording
    state = [10, 0, 0]
    assert _replay_exit_rule([1.0, 0.5, 0.25], 2e-7, 20, state) is None and state[2] == 3
    assert _replay_exit_rule([0.1, 1e-8, 0.0], 2e-7, 20, state) == 5


def _bc_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import torch.nn as nn
    import fakelib
    from dfq_b200 import dfq, workload, dist as ddist
    torch.set_num_threads(2)
    fakelib.install_plain()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        targ = [nn.Conv2d, nn.Linear]
        out = {}
        # (a) a serial network: every level replicated, nothing exchanged; (b) 6 independent blocks in ONE level, sharded
        for tag, topo, kw in (("mbv2", workload.load_topology(os.path.join(GOLD, "topology_mobilenetv2.json")), {}),
                              ("stack", workload.stack_topology(6, channels=48, k=3), dict(replicate_below=1))):
            def prepared():
                graph, bottoms, _ = workload.build_graph(topo, seed=4)
                for m in graph.values():
                    if isinstance(m, nn.BatchNorm2d):
                        m.register_buffer("fake_weight", m.weight.detach().abs().clone())
                        m.register_buffer("fake_bias", m.bias.detach().clone())
                return graph, bottoms
            ga, ba = prepared()
            dfq.bias_correction(ga, ba, targ)                      # what one process computes
            gb, bb = prepared()
            info = ddist.sharded_bias_correction(gb, bb, targ, **kw)
            out[tag + "_sharded_levels"] = np.array(info["sharded_levels"])
            out[tag + "_owners"] = np.array(sorted(set(info["owner"].values())))
            worst = 0.0
            for (ka, ma), (kb, mb) in zip(ga.items(), gb.items()):
                if type(ma) in targ and ma.bias is not None:
                    assert mb.bias is not None
                    worst = max(worst, _nw(mb.bias.detach().numpy(), ma.bias.detach().numpy()))
                    assert np.array_equal(ma.weight.detach().numpy(), mb.weight.detach().numpy())
                if hasattr(ma, "fake_bias") and not isinstance(ma, str):
                    worst = max(worst, _nw(mb.fake_bias.numpy(), ma.fake_bias.numpy()))
            out[tag + "_worst"] = np.array(worst)
        np.savez(os.path.join(out_dir, "bc%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_sharded_bias_correction(tmp_path):
    """sharded_bias_correction: both ranks end with the biases / fake_bias vectors a single process computes - bit for bit
    (the replicated levels run the same deterministic code, the sharded level copies the owner's rows)."""
    port = 29500 + (os.getpid() % 2000) + 41
    mp.spawn(_bc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        d = np.load(tmp_path / ("bc%d.npz" % r))
        assert float(d["mbv2_worst"]) == 0.0 and float(d["stack_worst"]) == 0.0, (r, float(d["mbv2_worst"]), float(d["stack_worst"]))
        assert int(d["mbv2_sharded_levels"]) == 0 and d["mbv2_owners"].tolist() == [-1]
        assert int(d["stack_sharded_levels"]) == 1 and d["stack_owners"].tolist() == [0, 1]


def _observer_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import torch.nn as nn
    import fakelib
    from dfq_b200 import dist as ddist
    from dfq_b200.utils.quantize import QuantMeasure
    torch.set_num_threads(1)
    fakelib.install_plain()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        model = nn.Sequential(QuantMeasure(True, 8), nn.ReLU(), nn.Sequential(QuantMeasure(True, 8), QuantMeasure(True, 8)))
        obs = [m for m in model.modules() if isinstance(m, QuantMeasure)]
        g = torch.Generator().manual_seed(100 + rank)          # what this rank's share of the batches left behind
        for m in obs:
            m.running_min.fill_(float(-torch.rand(1, generator=g)))
            m.running_max.fill_(float(torch.rand(1, generator=g)))
        before = np.array([[float(m.running_min), float(m.running_max)] for m in obs])
        n = ddist.sync_observers(model)
        after = np.array([[float(m.running_min), float(m.running_max)] for m in obs])
        np.savez(os.path.join(out_dir, "obs%d.npz" % rank), before=before, after=after, n=n)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_observer_sync(tmp_path):
    """sync_observers: every rank ends with min over ranks of running_min and max over ranks of running_max of every
    QuantMeasure (what one process would have reached over all batches in update_stat mode, quantize.py:103-107)."""
    port = 29500 + (os.getpid() % 2000) + 23
    mp.spawn(_observer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "obs0.npz"), np.load(tmp_path / "obs1.npz")
    assert int(r0["n"]) == 3 and int(r1["n"]) == 3
    want = np.stack([np.minimum(r0["before"][:, 0], r1["before"][:, 0]), np.maximum(r0["before"][:, 1], r1["before"][:, 1])], axis=1)
    assert np.array_equal(r0["after"], want) and np.array_equal(r1["after"], want)


class _TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        nn = torch.nn
        self.c1 = nn.Conv2d(3, 6, 3, stride=4, padding=1, bias=False); self.b1 = nn.BatchNorm2d(6)
        self.c2 = nn.Conv2d(6, 8, 3, stride=4, padding=1, bias=False); self.b2 = nn.BatchNorm2d(8)

    def forward(self, x):
        return self.b2(self.c2(torch.relu(self.b1(self.c1(x))))).mean((2, 3))


def _tiny_net():
    torch.manual_seed(3)
    m = _TinyNet().eval()
    for bn in (m.b1, m.b2):
        bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    return m


def _distill_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dfq_b200 import distill
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        torch.manual_seed(99)
        out = distill.getDistilData(_tiny_net(), "imagenet", 2, num_batch=3, gpu=False, value_range=[-2.5, 2.5], size=[224, 224],
                                    early_break_factor=1e-9, iterations=3)
        np.savez(os.path.join(out_dir, "distill%d.npz" % rank), **{"b%d" % i: t.numpy() for i, t in enumerate(out)})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_distilled_data_equals_one_process(tmp_path):
    """getDistilData (ZeroQ/distill_data.py:75-227) with a process group: the batches are independent optimisations dealt
    round-robin to the ranks; every rank draws the initial noise of ALL batches, so batch i starts where a single process
    would have started it, and after the final broadcasts every rank holds the list one process would have produced."""
    from dfq_b200 import distill
    torch.manual_seed(99)
    alone = distill.getDistilData(_tiny_net(), "imagenet", 2, num_batch=3, gpu=False, value_range=[-2.5, 2.5], size=[224, 224],
                                  early_break_factor=1e-9, iterations=3)
    port = 29500 + (os.getpid() % 2000) + 31
    mp.spawn(_distill_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in (0, 1):
        got = np.load(tmp_path / ("distill%d.npz" % rank))
        for i, t in enumerate(alone):
            assert np.array_equal(got["b%d" % i], t.numpy()), (rank, i)
