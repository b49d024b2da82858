"""-m gpu: the remaining public entry points of the path, real CUDA library vs the oracle on the same seeded inputs.

    _quantize_error      dfq.py:8-25          all five reductions, signed / unsigned, CPU and CUDA inputs
    bias_absorption      dfq.py:121-164       dense, pointwise -> depthwise (G = C) and depthwise -> pointwise second layers
    clip_weight          dfq.py:167-170
    QConv2d / QLinear / QuantLinear forward   quantize.py:124-205,253-341  incl. merge_scale_prev / merge_scale
    update_quant_range   improve_dfq.py:280-297
    dfq_range_rows / dfq_range_cols           the stand-alone per-channel range entry points of the C ABI
    dfq_bias_correct alone on REFERENCE-produced post-equalization weights (full tensors), 1e-5
"""
import ctypes as C
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import dfq_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def _nw(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("device", ["cpu", "cuda"])
@pytest.mark.parametrize("signed", [False, True])
def test_quantize_error_every_reduction(device, signed):
    """dfq.py:8-25.  The element-wise error (reduction=None) is bit-exact: CPU inputs follow PyTorch-CPU's true division,
    CUDA inputs the reciprocal multiply PyTorch-CUDA performs for div_(python_float).  The reductions are torch's own on
    that tensor; they are compared with float64 reductions of the oracle's error at 1e-5."""
    from dfq_b200.dfq import _quantize_error
    g = torch.Generator().manual_seed(11)
    for shape in ((16, 8, 3, 3), (32, 1, 3, 3), (24, 96, 1, 1), (10, 128)):
        w = torch.randn(*shape, generator=g) * 0.7
        w.view(-1)[5] = 3.1
        keep = w.clone()
        x = w.to(device)
        ref = O.quantize(w.numpy(), 8, float(w.min()), float(w.max()), signed, div_mode="div" if device == "cpu" else "recip") - w.numpy()
        got = _quantize_error(x, 8, None, signed)
        assert got.device.type == device and got.shape == w.shape
        assert np.array_equal(got.cpu().numpy(), ref), (shape, device, signed)
        assert torch.equal(x.cpu(), keep), "_quantize_error must not touch its argument (it clones, dfq.py:12)"
        r64 = ref.astype(np.float64)
        want = {"sum": np.abs(r64).sum(), "mean": r64.mean(),
                "channel": np.abs(r64.reshape(shape[0], -1).sum(-1)).sum()}
        if len(shape) == 4:
            want["spatial"] = np.abs(r64.reshape(shape[0], shape[1], -1).sum(-1)).sum()
        for red, val in want.items():
            out = _quantize_error(x, 8, red, signed)
            assert out.dim() == 0 and out.device.type == device
            tol = 1e-5 * max(abs(val), np.abs(r64).sum() / r64.size if red == "mean" else 0.0) + 1e-9
            assert abs(float(out) - val) <= tol, (shape, red, float(out), val)
    # other bit widths take the same path
    w = torch.randn(64, 32, generator=g)
    for bits in (4, 16):
        ref = O.quantize(w.numpy(), bits, float(w.min()), float(w.max()), signed) - w.numpy()
        assert np.array_equal(_quantize_error(w, bits, None, signed).numpy(), ref)


# ---------------------------------------------------------------------------------------------------------------------
def _absorption_graph(seed, shapes, device):
    """conv -> BN -> ReLU -> conv -> BN -> ReLU -> ... as plain graph / bottoms dictionaries."""
    g = torch.Generator().manual_seed(seed)
    graph, bottoms = OrderedDict(Data="Data"), OrderedDict(Data=None)
    prev = "Data"
    convs, bns = [], []
    for i, (o, j, k, groups, has_bias) in enumerate(shapes):
        conv = nn.Conv2d(j * groups, o, k, groups=groups, bias=has_bias)
        conv.weight.data = torch.randn(o, j, k, k, generator=g) * 0.3
        if has_bias:
            conv.bias.data = torch.randn(o, generator=g)
        bn = nn.BatchNorm2d(o)
        bn.register_buffer("fake_weight", torch.rand(o, generator=g) * 0.4 + 0.05)
        bn.register_buffer("fake_bias", torch.randn(o, generator=g) * 1.5)        # beta - 3*gamma > 0 for a good share
        relu = nn.ReLU()
        conv.to(device); bn.to(device)
        for key, mod in ((id(conv), conv), (id(bn), bn), (id(relu), relu)):
            graph[key] = mod
            bottoms[key] = [prev]
            prev = key
        convs.append(conv); bns.append(bn)
    return graph, bottoms, convs, bns


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_bias_absorption_matches_oracle(device):
    """dfq.py:121-164 on a five-layer chain: dense 3x3 -> pointwise -> depthwise (second layer with G = C groups) ->
    pointwise (depthwise as FIRST) -> dense, with and without existing biases.  c, the two `-= c` updates and fake_bias
    are exact fp32 ops (bit-exact); wc = (sum_k W2) @ c is a reduction (1e-5 normwise)."""
    from dfq_b200.dfq import bias_absorption
    from dfq_b200.utils.relation import Relation
    shapes = [(16, 8, 3, 1, True), (24, 16, 1, 1, False), (24, 1, 3, 24, True), (12, 24, 1, 1, False), (20, 12, 3, 1, True)]
    graph, bottoms, convs, bns = _absorption_graph(5, shapes, device)
    rels = [Relation(id(convs[i]), id(convs[i + 1]), id(bns[i])) for i in range(4)]
    w0 = [c.weight.detach().cpu().numpy().copy() for c in convs]
    b0 = [None if c.bias is None else c.bias.detach().cpu().numpy().copy() for c in convs]
    fw0 = [b.fake_weight.cpu().numpy().copy() for b in bns]
    fb0 = [b.fake_bias.cpu().numpy().copy() for b in bns]
    bias_absorption(graph, rels, bottoms, 3)
    # oracle, in the reference's order (dfq.py:162-164 per relation)
    b_ref = [np.zeros(w.shape[0], f32) if b is None else b.copy() for w, b in zip(w0, b0)]
    fb_ref = [v.copy() for v in fb0]
    some_positive = False
    for i in range(4):
        c = O.bias_absorb_c(fw0[i], fb_ref[i], 3)
        some_positive |= bool((c > 0).any())
        wc = O.bias_absorb_wc(w0[i + 1], c, w0[i].shape[0])
        b_ref[i] = b_ref[i] + (-c)
        fb_ref[i] = fb_ref[i] + (-c)
        b_ref[i + 1] = b_ref[i + 1] + wc
    assert some_positive
    for i, conv in enumerate(convs):
        assert np.array_equal(conv.weight.detach().cpu().numpy(), w0[i]), "absorption must not touch weights"
        assert conv.bias is not None and conv.bias.device.type == device
        assert _nw(conv.bias.detach().cpu().numpy(), b_ref[i]) < 1e-5, (i, _nw(conv.bias.detach().cpu().numpy(), b_ref[i]))
    for i in range(4):
        assert np.array_equal(bns[i].fake_bias.cpu().numpy(), fb_ref[i]), i
        assert np.array_equal(bns[i].fake_weight.cpu().numpy(), fw0[i])
    # the last layer of the chain is only ever `second`: bias = b0 + wc, nothing subtracted
    assert np.array_equal(bns[4].fake_bias.cpu().numpy(), fb0[4])


def test_bias_absorption_skips_relations_without_relu_and_handles_any_order():
    from dfq_b200.dfq import bias_absorption
    from dfq_b200.utils.relation import Relation
    shapes = [(8, 4, 3, 1, True), (8, 8, 1, 1, True), (6, 8, 1, 1, True)]
    graph, bottoms, convs, bns = _absorption_graph(9, shapes, "cpu")
    # drop the ReLU between conv0/bn0 and conv1: relation 0 has no ReLU on its path -> skipped (dfq.py:136-137)
    relu0 = [k for k in graph if isinstance(graph[k], nn.ReLU)][0]
    nxt = [k for k in graph if bottoms[k] == [relu0]][0]
    bottoms[nxt] = bottoms[relu0]
    del graph[relu0], bottoms[relu0]
    rels = [Relation(id(convs[1]), id(convs[2]), id(bns[1])), Relation(id(convs[0]), id(convs[1]), id(bns[0]))]   # backward order
    b0 = [c.bias.detach().numpy().copy() for c in convs]
    fb0 = [b.fake_bias.numpy().copy() for b in bns]
    bias_absorption(graph, rels, bottoms, 3)
    c = O.bias_absorb_c(bns[1].fake_weight.numpy(), fb0[1], 3)
    wc = O.bias_absorb_wc(convs[2].weight.detach().numpy(), c, 8)
    assert np.array_equal(convs[0].bias.detach().numpy(), b0[0]) and np.array_equal(bns[0].fake_bias.numpy(), fb0[0])
    assert np.array_equal(convs[1].bias.detach().numpy(), b0[1] + (-c))
    assert _nw(convs[2].bias.detach().numpy(), b0[2] + wc) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_clip_weight_matches_oracle(device):
    from dfq_b200.dfq import clip_weight
    g = torch.Generator().manual_seed(2)
    conv = nn.Conv2d(8, 16, 3).to(device); lin = nn.Linear(33, 7).to(device); other = nn.ConvTranspose2d(4, 4, 3).to(device)
    for m in (conv, lin, other):
        m.weight.data = (torch.randn(m.weight.shape, generator=g) * 12).to(device)
    graph = OrderedDict(a=conv, b="add_3", c=lin, d=other)
    before = [m.weight.detach().cpu().numpy().copy() for m in (conv, lin, other)]
    ptrs = [m.weight.data_ptr() for m in (conv, lin)]
    clip_weight(graph, range_clip=[-15, 15], targ_type=[nn.Conv2d, nn.Linear])
    assert np.array_equal(conv.weight.detach().cpu().numpy(), O.clip_weight(before[0]))
    assert np.array_equal(lin.weight.detach().cpu().numpy(), O.clip_weight(before[1]))
    assert np.array_equal(other.weight.detach().cpu().numpy(), before[2]), "types outside targ_type are left alone"
    assert (np.abs(before[0]) > 15).any() and [m.weight.data_ptr() for m in (conv, lin)] == ptrs
    clip_weight(graph, range_clip=[-0.5, 2.0], targ_type=[nn.Linear])
    assert np.array_equal(lin.weight.detach().cpu().numpy(), O.clip_weight(before[1], -0.5, 2.0))


# ---------------------------------------------------------------------------------------------------------------------
def _eager_q(x, bits, mn, mx):
    """The reference's op chain (quantize.py:70-74) executed by PyTorch itself on x's device, explicit float range."""
    qmax = 2. ** bits - 1.
    scale = max((mx - mn) / qmax, 1e-8)
    return x.clone().add_(-mn).div_(scale).clamp_(0., qmax).round_().mul_(scale).add_(mn)


def _eager_q_implicit(b, bits):
    """min_value=None path (quantize.py:24-35,64-66): 0-d fp32 tensor range, tensor division."""
    y = b.reshape(1, -1)
    mn = y.min(-1)[0].mean(-1); mx = y.max(-1)[0].mean(-1)
    qmax = 2. ** bits - 1.
    scale = (mx - mn) / qmax
    scale = torch.clamp(scale, min=1e-8) if isinstance(scale, torch.Tensor) else max(scale, 1e-8)
    return b.clone().add_(-mn).div_(scale).clamp_(0., qmax).round_().mul_(scale).add_(mn)


def test_q_layers_forward_with_learned_scales_match_eager_chain():
    """QConv2d / QLinear / QuantLinear on the GPU (quantize.py:124-205,253-341): the input observer, merge_scale_prev per
    group (a division for convs, quantize.py:158-167; a product for linears, :283), merge_scale, per-forward weight
    quantization with the tensor's own range and implicit-range bias quantization - against the same op chain run by
    PyTorch CUDA eager.  Quantized operands must be bit-equal; the convolution itself is cuDNN on both sides."""
    from dfq_b200.utils import quantize as Q
    torch.manual_seed(4)
    x = torch.randn(4, 12, 10, 10, device="cuda") * 1.5
    for groups in (1, 3):
        conv = Q.QConv2d(12, 18, 3, padding=1, groups=groups, num_bits=8, num_bits_act=8, num_bits_bias=16).cuda().eval()
        conv.quant.running_min.fill_(-4.); conv.quant.running_max.fill_(4.5)
        scale = torch.rand(18, device="cuda") + 0.5
        scale_prev = (torch.rand(12, device="cuda") + 0.5).view(-1, 1, 1, 1)
        conv.set_scale(scale=scale.clone(), scale_prev=scale_prev.clone())
        y = conv(x)
        w = conv.weight.detach()
        rows, cols = 18 // groups, 12 // groups
        sw = w.clone()
        for gi in range(groups):
            sw[gi * rows:(gi + 1) * rows] = w[gi * rows:(gi + 1) * rows] / scale_prev[:, 0, 0, 0].view(1, -1, 1, 1)[:, gi * cols:(gi + 1) * cols]
        sw = sw * scale.view(-1, 1, 1, 1)
        sb = conv.bias.detach() * scale
        qw = _eager_q(sw, 8, float(sw.min()), float(sw.max()))
        qb = _eager_q_implicit(sb, 16)
        qx = _eager_q(x, 8, -4., 4.5)
        got_w = Q._quant_param_per_forward(sw, 8)
        assert torch.equal(got_w, qw), (groups, (got_w != qw).sum().item())
        assert torch.equal(Q.quantize(sb, num_bits=16), qb)
        assert torch.equal(conv.quant(x), qx)
        ref = torch.nn.functional.conv2d(qx, qw, qb, 1, 1, 1, groups)
        assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5)
        # merge_scale_to_weight folds both scales into the parameters; the forward result stays the same
        conv.merge_scale_to_weight()
        assert getattr(conv, "scale", None) is None and getattr(conv, "scale_prev", None) is None
        assert torch.allclose(conv.weight.detach(), sw, rtol=0, atol=0)
        assert torch.allclose(conv(x), ref, rtol=1e-5, atol=1e-5)
    xl = torch.randn(6, 40, device="cuda")
    lin = Q.QLinear(40, 24, num_bits=8, num_bits_act=8, num_bits_bias=16).cuda().eval()
    lin.quant.running_min.fill_(-3.); lin.quant.running_max.fill_(3.)
    s, sp = torch.rand(24, device="cuda") + 0.5, torch.rand(40, device="cuda") + 0.5
    lin.set_scale(scale=s.clone(), scale_prev=sp.clone())
    sw = lin.weight.detach() * sp.view(1, -1) * s.view(-1, 1)
    sb = lin.bias.detach() * s
    ref = torch.nn.functional.linear(_eager_q(xl, 8, -3., 3.), _eager_q(sw, 8, float(sw.min()), float(sw.max())), _eager_q_implicit(sb, 16))
    assert torch.allclose(lin(xl), ref, rtol=1e-5, atol=1e-5)
    ql = Q.QuantLinear(40, 24, num_bits=4, num_bits_act=8, num_bits_bias=8).cuda().eval()
    ql.quant.running_min.fill_(-3.); ql.quant.running_max.fill_(3.)
    w = ql.weight.detach()
    qw4 = _eager_q(w, 4, float(w.min()), float(w.max()))
    assert torch.equal(Q._quant_param_per_forward(w, 4), qw4) and len(torch.unique(qw4)) <= 16
    ref = torch.nn.functional.linear(_eager_q(xl, 8, -3., 3.), qw4, _eager_q_implicit(ql.bias.detach(), 8))
    assert torch.allclose(ql(xl), ref, rtol=1e-5, atol=1e-5)
    # straight-through gradient (quantize.py:78-83)
    xg = xl.clone().requires_grad_(True)
    ql(xg).sum().backward()
    assert xg.grad is not None and ql.weight.grad is not None and torch.isfinite(xg.grad).all()


# ---------------------------------------------------------------------------------------------------------------------
def test_update_quant_range_drives_every_observer_and_pins_the_input_range():
    """improve_dfq.py:280-297 on a small model whose forward adds two branches: observers in update_stat mode follow
    quantize.py:103-107 over the batches, the functional-op observers fire through replace_op(), the layer fed by 'Data'
    gets the preprocessing constants, and the patched ops are restored afterwards."""
    from dfq_b200.improve_dfq import set_update_stat, update_quant_range
    from dfq_b200.utils import layer_transform as LT
    from dfq_b200.utils.quantize import QuantMeasure, QuantNConv2d

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1 = QuantNConv2d(3, 8, 3, padding=1)
            self.c2 = QuantNConv2d(8, 8, 3, padding=1)

        def forward(self, x):
            a = self.c1(x)
            b = self.c2(a)
            return a + b
    torch.manual_seed(0)
    model = Net().eval()
    add_line = Net.forward.__code__.co_firstlineno + 3
    record = [("add_9", "add_%d_2" % add_line)]
    ops = [QuantMeasure(num_bits=8, momentum=0.1) for _ in range(2)]
    LT.module_tensor_op = LT.CustomTensorOP(ops, record)
    model.add_module("custom_tensor_op", LT.module_tensor_op)
    model.eval()     # (in main_cls.py the op observers are added after model.eval() and stay in training mode - EMA on top,
    #                   quantize.py:109-113 - until the final model.eval(); here every observer is in eval mode)
    graph = OrderedDict([("Data", "Data"), (1, model.c1), (2, model.c2), ("add_9", "add_9")])
    bottoms = OrderedDict([("Data", None), (1, ["Data"]), (2, [1]), ("add_9", [1, 2])])
    g = torch.Generator().manual_seed(3)
    data = [torch.randn(8, 3, 16, 16, generator=g).clamp_(-2.1, 2.6) for _ in range(3)]
    raw_add = torch.Tensor.__add__
    set_update_stat(model, [QuantMeasure], True)
    model = update_quant_range(model.cuda(), data, graph, bottoms)
    set_update_stat(model, [QuantMeasure], False)
    assert torch.Tensor.__add__ is raw_add, "restore_op must put the original operators back"
    assert abs(float(model.c1.quant.running_min) + 2.11790393) < 1e-6 and abs(float(model.c1.quant.running_max) - 2.64) < 1e-6
    # replay with plain torch: what each observer saw and the statistic it must hold
    want = {"c2": [0., 0.], "op0": [0., 0.], "op1": [0., 0.]}

    def upd(key, t):
        mn, mx = O.per_sample_minmax_mean(t.reshape(t.shape[0], -1).cpu().numpy())
        want[key] = [min(want[key][0], float(mn)), max(want[key][1], float(mx))]
    rmin, rmax = -2.11790393, 2.64        # c1's range only matters for what c1 passes on: its observer is in update mode too
    c1_min, c1_max = 0., 0.
    with torch.no_grad():
        for x in data:
            xc = x.cuda()
            mn, mx = O.per_sample_minmax_mean(x.reshape(8, -1).numpy())
            c1_min, c1_max = min(c1_min, float(mn)), max(c1_max, float(mx))
            qx = torch.from_numpy(O.quantize(x.numpy(), 8, c1_min, c1_max, div_mode="recip")).cuda()
            a = torch.nn.functional.conv2d(qx, model.c1.weight, model.c1.bias, 1, 1)
            upd("c2", a)
            qa = torch.from_numpy(O.quantize(a.cpu().numpy(), 8, want["c2"][0], want["c2"][1], div_mode="recip")).cuda()
            b = torch.nn.functional.conv2d(qa, model.c2.weight, model.c2.bias, 1, 1)
            upd("op0", a); upd("op1", b)
    for key, qm in (("c2", model.c2.quant), ("op0", ops[0]), ("op1", ops[1])):
        got = [float(qm.running_min), float(qm.running_max)]
        assert np.allclose(got, want[key], rtol=2e-5, atol=1e-6), (key, got, want[key])
    assert not model.c2.quant.update_stat


# ---------------------------------------------------------------------------------------------------------------------
def test_range_rows_and_cols_entry_points():
    """dfq_range_rows / dfq_range_cols (dfq.py:50-55 as stand-alone C-ABI calls): exact min/max."""
    from dfq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(8)
    for (o, j, kk, groups) in ((32, 16, 9, 1), (24, 1, 9, 24), (40, 24, 1, 1), (1000, 1280, 1, 1), (36, 12, 9, 3), (5, 7, 25, 1)):
        w = torch.randn(o, j, kk, generator=g).cuda().contiguous()
        rmin = torch.empty(o, device="cuda"); rmax = torch.empty(o, device="cuda")
        _lib.check(lib.dfq_range_rows(C.c_void_p(w.data_ptr()), o, j * kk, C.c_void_p(rmin.data_ptr()), C.c_void_p(rmax.data_ptr()),
                                      _lib.stream_ptr()), "dfq_range_rows")
        assert torch.equal(rmin, w.view(o, -1).min(-1)[0]) and torch.equal(rmax, w.view(o, -1).max(-1)[0])
        cmin = torch.empty(groups * j, device="cuda"); cmax = torch.empty(groups * j, device="cuda")
        _lib.check(lib.dfq_range_cols(C.c_void_p(w.data_ptr()), o, j, kk, groups, C.c_void_p(cmin.data_ptr()),
                                      C.c_void_p(cmax.data_ptr()), _lib.stream_ptr()), "dfq_range_cols")
        v = w.view(groups, o // groups, j, kk)
        assert torch.equal(cmin, v.amin(dim=(1, 3)).reshape(-1)) and torch.equal(cmax, v.amax(dim=(1, 3)).reshape(-1)), (o, j, kk, groups)
    assert lib.dfq_range_rows(None, 4, 4, None, None, None) != 0 and b"bad" in lib.dfq_last_error()


# ---------------------------------------------------------------------------------------------------------------------
def run_bias_correction_against_reference_fixture(name):
    """Shared with the CPU twin in tests/test_host_logic.py (oracle-backed executor)."""
    import bc_fixture
    from dfq_b200 import dfq
    gold = np.load(os.path.join(GOLD, "ref_bc_%s.npz" % name))
    graph, bottoms = bc_fixture.build(name)
    assert np.array_equal(bc_fixture.input_digests(graph), gold["digests"]), \
        "regenerated inputs differ from the ones the reference saw (torch RNG changed?) - regenerate with tools/make_golden.py bc"
    w0 = {k: m.weight.detach().numpy().copy() for k, m in graph.items() if type(m) in bc_fixture.TARG}
    dfq.bias_correction(graph, bottoms, bc_fixture.TARG)
    worst, n = 0.0, 0
    for i, k in enumerate(graph):
        m = graph[k]
        if type(m) in bc_fixture.TARG:
            assert np.array_equal(m.weight.detach().numpy(), w0[k]), "bias correction must not touch weights"
            if "out_b_%d" % i in gold.files:
                e = _nw(m.bias.detach().numpy(), gold["out_b_%d" % i]); worst = max(worst, e); n += 1
                assert e < 1e-5, ("bias", i, e)
        elif "out_fb_%d" % i in gold.files:
            e = _nw(m.fake_bias.numpy(), gold["out_fb_%d" % i]); worst = max(worst, e)
            assert e < 1e-5, ("fake_bias", i, e)
    assert n == len([k for k in gold.files if k.startswith("out_b_")])
    return worst


@pytest.mark.parametrize("name", ["resnet18", "mobilenetv2"])
def test_bias_correction_alone_against_reference_produced_numbers(name):
    """dfq_bias_correct vs numbers the REFERENCE produced (tests/golden/ref_bc_<model>.npz, tools/make_golden.py bc): the
    inputs are regenerated bit-identically from a seed (tests/bc_fixture.py, sha-checked), the expected post-correction
    biases and fake_bias vectors come from running dfq.py:173-293 of the reference on them, so nothing but bias correction
    sits between input and expected output (no dependence on the equalization's last bit).
    Gate: 1e-5 normwise."""
    worst = run_bias_correction_against_reference_fixture(name)
    print("bias correction vs reference-produced numbers (%s): worst normwise error %.3g" % (name, worst))


# ---------------------------------------------------------------------------------------------------------------------
def test_fused_observer_matches_reference_vectors_and_is_one_launch():
    """QuantMeasure.forward as ONE launch (dfq_observe_quant): update_stat / training-EMA / eval paths against the vectors
    the REFERENCE's QuantMeasure produced on the CPU (tests/golden/ref_ops.npz: obs_*, ema_*), plus the own-range modes
    (per-forward weight quantization, implicit-range bias quantization qimp_*) and a large activation tensor against the
    oracle; every case is also compared with the separate launches (statistic, update, quantize) bit for bit."""
    import ctypes as C
    from dfq_b200 import _lib
    from dfq_b200.utils import quantize as Q
    ops = np.load(os.path.join(GOLD, "ref_ops.npz"))
    # (1) update_stat in eval mode: running range follows the batch statistic, output quantized with the UPDATED range
    x = torch.from_numpy(ops["obs_in"].copy())
    qm = Q.QuantMeasure(True).eval()
    y = qm(x)                                   # CPU input: staged through the GPU, true division like the reference's CPU run
    assert abs(float(qm.running_min) - float(ops["obs_min"])) <= 1e-6 * abs(float(ops["obs_min"]))
    assert abs(float(qm.running_max) - float(ops["obs_max"])) <= 1e-6 * abs(float(ops["obs_max"]))
    want = O.quantize(ops["obs_in"], 8, float(qm.running_min), float(qm.running_max))
    assert np.array_equal(y.numpy(), want)
    assert np.abs(y.numpy() - ops["obs_out"]).max() <= 1.001 * (float(ops["obs_max"]) - float(ops["obs_min"])) / 255
    # (2) training: EMA of the statistic, quantized with the statistic itself
    qm2 = Q.QuantMeasure(False).train()
    y2 = qm2(x)
    assert abs(float(qm2.running_min) - float(ops["ema_min"])) <= 2e-6 * abs(float(ops["ema_min"]))
    assert abs(float(qm2.running_max) - float(ops["ema_max"])) <= 2e-6 * abs(float(ops["ema_max"]))
    mn, mx = O.per_sample_minmax_mean(ops["obs_in"].reshape(ops["obs_in"].shape[0], -1))
    assert np.array_equal(y2.detach().numpy(), O.quantize(ops["obs_in"], 8, float(mn), float(mx)))
    # (3) implicit-range (bias) path: bit-exact against the reference's vectors
    b = torch.from_numpy(ops["qimp_in"].copy())
    assert np.array_equal(Q.quantize(b, num_bits=16).numpy(), ops["qimp_out16"])
    assert np.array_equal(Q.quantize(b, num_bits=8).numpy(), ops["qimp_out8"])
    # (4) a large CUDA activation, all three flag combinations, against the separate launches
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    xa = torch.randn(64, 64, 56, 56, device="cuda", generator=g) * 1.7          # 12.8 M elements (ResNet-18 conv input)
    P = lambda t: C.c_void_p(t.data_ptr())
    for flags in (1, 2, 3):
        rmin = torch.tensor([-0.5], device="cuda"); rmax = torch.tensor([0.7], device="cuda")
        fused = Q.observe_and_quant(xa, 8, flags, rmin, rmax, 0.1)
        smin = torch.tensor([-0.5], device="cuda"); smax = torch.tensor([0.7], device="cuda")
        stat = Q.per_sample_minmax_mean(xa)
        if flags & 1:
            _lib.check(lib.dfq_observer_update(P(smin), P(smax), P(stat), 1, C.c_float(0.1), _lib.stream_ptr()), "upd")
        if flags & 2:
            _lib.check(lib.dfq_observer_update(P(smin), P(smax), P(stat), 2, C.c_float(0.1), _lib.stream_ptr()), "ema")
            sep = Q.fake_quant_device_range(xa, 8, stat[0:1], stat[1:2])
        else:
            sep = Q.fake_quant_device_range(xa, 8, smin, smax)
        assert torch.equal(rmin, smin) and torch.equal(rmax, smax), flags
        assert torch.equal(fused, sep), (flags, (fused != sep).sum().item())
    # per-sample statistic itself vs plain torch
    want_max = xa.view(64, -1).max(-1)[0].double().mean(); want_min = xa.view(64, -1).min(-1)[0].double().mean()
    rmin = torch.zeros(1, device="cuda"); rmax = torch.zeros(1, device="cuda")
    Q.observe_and_quant(xa, 8, 1, rmin, rmax, 0.1)
    assert abs(float(rmax) - float(want_max)) <= 1e-6 * float(want_max) and abs(float(rmin) - float(want_min)) <= 1e-6 * abs(float(want_min))
    # (5) own range, batch 1: quantize(w, bits, float(w.min()), float(w.max())) on a CUDA weight
    w = torch.randn(256, 128, 3, 3, device="cuda", generator=g)
    got = Q._quant_param_per_forward(w, 8)
    assert torch.equal(got, _eager_q(w, 8, float(w.min()), float(w.max())))
