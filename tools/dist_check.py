"""Multi-GPU parity check: the chains of a model are sharded over the ranks (one process per
GPU, NCCL), every rank equalizes its chains with the CUDA engine, ONE all-gather of the scale vectors closes the step and
the other ranks' chains are replayed from the gathered scales.  Every rank must end with the model a single-GPU run gives:
S bit for bit, weights within 1e-5 (they are bit-identical in practice).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/dist_check.py [deeplab|mobilenetv2|resnet18|ssd] [exact|per_chain]
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from dfq_b200 import dfq, workload, dist as ddist
from dfq_b200.utils import layer_transform as LT
from dfq_b200.utils.relation import create_relation


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "deeplab"
    mode = sys.argv[2] if len(sys.argv) > 2 else "exact"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    targ = [nn.Conv2d, nn.Linear]
    topo = workload.load_topology(os.path.join(ROOT, "tests", "golden", "topology_%s.json" % name))

    def prepared():
        graph, bottoms, _ = workload.build_graph(topo, seed=0)
        LT.merge_batchnorm(None, graph, bottoms, targ)
        return graph, create_relation(graph, bottoms, targ)

    ga, ra = prepared()                       # single GPU (every rank computes it: the reference result)
    dfq.cross_layer_equalization(ga, ra, targ)
    gb, rb = prepared()                       # sharded
    info = ddist.sharded_cross_layer_equalization(gb, rb, targ, mode=mode)
    worst, s_equal = 0.0, True
    for x, y in zip(ra, rb):
        s_equal &= bool(torch.equal(x.S, y.S))
    for (ka, ma), (kb, mb) in zip(ga.items(), gb.items()):
        if type(ma) in targ:
            a, b = ma.weight.detach().double(), mb.weight.detach().double()
            worst = max(worst, float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)))
    # bias correction: single process on the single-GPU result vs sharded_bias_correction on the sharded result (both models are
    # equal at this point up to the replay tolerance above; the correction itself must agree to 1e-5)
    ga2, ba2, _ = workload.build_graph(topo, seed=0)
    LT.merge_batchnorm(None, ga2, ba2, targ)
    ra2 = create_relation(ga2, ba2, targ)
    dfq.cross_layer_equalization(ga2, ra2, targ)
    gb2, bb2, _ = workload.build_graph(topo, seed=0)
    LT.merge_batchnorm(None, gb2, bb2, targ)
    rb2 = create_relation(gb2, bb2, targ)
    dfq.cross_layer_equalization(gb2, rb2, targ)               # identical starting points for the correction
    dfq.bias_correction(ga2, ba2, targ)
    binfo = ddist.sharded_bias_correction(gb2, bb2, targ, replicate_below=1 << 18)
    worst_b = 0.0
    for (ka, ma), (kb, mb) in zip(ga2.items(), gb2.items()):
        if type(ma) in targ and ma.bias is not None:
            a, b = ma.bias.detach().double(), mb.bias.detach().double()
            worst_b = max(worst_b, float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)))
        if hasattr(ma, "fake_bias") and not isinstance(ma, str):
            a, b = ma.fake_bias.double(), mb.fake_bias.double()
            worst_b = max(worst_b, float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)))
    flag = torch.tensor([1.0 if (s_equal or mode != "exact") and worst <= 1e-5 and worst_b <= 1e-5 else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        owners = np.bincount(np.asarray(info["owner"]), minlength=world).tolist()
        print("dist_check %s mode=%s world=%d: chains per rank %s, sweeps %s, S bit-identical %s, worst weight normwise %.3g; "
              "bias correction: %d levels (%d sharded), worst bias/fake_bias normwise %.3g -> %s"
              % (name, mode, world, owners, info["sweeps"], s_equal, worst, binfo["levels"], binfo["sharded_levels"], worst_b,
                 "OK" if flag.item() == 1.0 else "MISMATCH"))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
