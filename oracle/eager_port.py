"""PyTorch-eager CPU port of the reference's calibration path WITH ITS EXECUTION STRUCTURE  -  TEST / BASELINE
INFRASTRUCTURE, NOT PRODUCT CODE (only tests/ and bench.py's CPU arm import it).

`oracle/dfq_oracle.py` is the vectorised numpy checker.  This file is the *timing* companion: it performs the same
arithmetic the way jakc4103/DFQ executes it on the host - one Python iteration per channel issuing a handful of tiny
eager tensor ops (dfq.py:48-73), a deep copy of all weights per sweep for the convergence test (dfq.py:84,105-108),
seven element-wise passes per fake quantization (utils/quantize.py:70-74) - because that execution structure, not
the arithmetic, is what the reference's CPU time consists of (~120 us per channel iteration).
It is validated against the numpy oracle in tests/test_oracle_pins.py.
"""
import torch


def equalize_pair_(w_a, w_b, b_a, bn_w, bn_b, lo=1e-8, hi=1e8, signed=False, eps=0):
    """Channel-at-a-time equalization of one relation (dfq.py:28-75); tensors are modified in place."""
    n_a, n_in_b = w_a.shape[0], w_b.shape[1]
    groups = 1 if n_a == n_in_b else n_a // n_in_b
    per_a, per_b = n_a // groups, w_b.shape[0] // groups
    scales = torch.zeros(n_a)
    for g in range(groups):
        rows_b = slice(g * per_b, (g + 1) * per_b)
        for j in range(n_in_b):
            c = g * per_a + j
            row, col = w_a[c], w_b[rows_b, j]
            if signed:
                r_a, r_b = row.abs().max(), col.abs().max()
            else:
                r_a, r_b = row.max() - row.min(), col.max() - col.min()
            s = (1 / (r_a + eps)) * torch.sqrt(r_a * r_b + eps)
            s = max(lo, min(hi, s))
            scales[c] = s
            row.mul_(s)
            for vec in (bn_w, bn_b, b_a):
                if vec is not None:
                    vec[c].mul_(s)
            col.mul_(1 / s)
    return scales


def fake_quant(x, bits=8):
    """utils/quantize.py:47-74 with the tensor's own range: clone + six in-place passes."""
    lo, hi = float(x.min()), float(x.max())
    qmax = 2. ** bits - 1.
    scale = max((hi - lo) / qmax, 1e-8)
    y = x.clone()
    y.add_(-lo).div_(scale)
    y.clamp_(0., qmax).round_()
    y.mul_(scale).add_(lo)
    return y


def relu_mean(gamma, beta):
    """dfq.py:182-184, 239-240 (scipy float64 pdf/cdf on fp32 arguments)."""
    from scipy.stats import norm
    q = -beta / gamma
    e = gamma * torch.from_numpy(norm(0, 1).pdf(q)).float() + beta * (1 - torch.from_numpy(norm.cdf(q)).float())
    e[e < 0] = 0
    return e


def calibrate_blocks(blocks, thres=2e-7, patience=20):
    """blocks: [[(w1, [gamma, beta, mean, var]), (w2, [...])], ...]  ->  number of equalization sweeps.

    BN fold (layer_transform.py:246-265), equalization with the reference's exit rule (dfq.py:81-115), bias correction of
    the second conv of every block (dfq.py:216-219, 281-293)."""
    with torch.no_grad():
        state = []
        for blk in blocks:
            layer = []
            for w, (gamma, beta, mean, var) in blk:
                den = torch.sqrt(var + 1e-5)
                w.mul_((gamma / den).view(-1, 1, 1, 1))
                bias = torch.zeros(w.shape[0]).mul(gamma / den).add(beta - (gamma * mean) / den)
                layer.append(dict(w=w, b=bias, fw=gamma.abs().clone(), fb=beta.clone()))
            state.append(layer)
        diff, count, sweeps = 10, 0, 0
        while diff > thres and count < patience:
            before = [[l["w"].clone() for l in layer] for layer in state]
            for layer in state:
                a, b = layer
                equalize_pair_(a["w"], b["w"], a["b"], a["fw"], a["fb"])
            cur = 0
            for layer, old in zip(state, before):
                for l, o in zip(layer, old):
                    cur += float(torch.mean(torch.abs(l["w"] - o)))
            sweeps += 1
            if abs(diff - cur) > 1e-9:
                count, diff = 0, cur
            else:
                count += 1
        for layer in state:
            a, b = layer
            err = (fake_quant(b["w"]) - b["w"]).view(b["w"].shape[0], b["w"].shape[1], -1).sum(-1)
            delta = torch.matmul(err, relu_mean(a["fw"], a["fb"]))
            b["b"].add_(-delta)
            b["fb"].add_(-delta)
        return sweeps
