"""Where the MobileNetV2 end-to-end milliseconds go (host gather / H2D / launches / D2H / host scatter).
usage: python tools/_mbv2_e2e_profile.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn as nn
from dfq_b200 import workload
from dfq_b200.calibrate import GraphCalibration

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
topo = workload.load_topology(os.path.join(root, "tests", "golden", "topology_mobilenetv2.json"))
graph, bottoms, modules = workload.build_graph(topo, seed=0)
backup = [{k: v.clone() for k, v in m.state_dict().items()} for m in modules]
cal = GraphCalibration(graph, bottoms, [nn.Conv2d, nn.Linear], device=torch.device("cuda:0"))
sess = cal.sess

eps0 = [getattr(m, "eps", None) for m in modules]

def restore():
    with torch.no_grad():
        for m, sd, e in zip(modules, backup, eps0):
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                m.weight.copy_(sd["weight"])
                if m.bias is not None:
                    m.bias.copy_(sd["bias"]) if "bias" in sd else m.bias.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                for k in ("weight", "bias", "running_mean", "running_var"):
                    getattr(m, k).copy_(sd[k])
                m.eps = e

T = time.perf_counter
sync = torch.cuda.synchronize
import ctypes as C
import numpy as np
from dfq_b200 import _lib
sess._ensure_room()
x = sess._transfer_lists()
lo, st = x["lo"], sess._staging
print("tensors up", len(x["h2d_bounds"]), "down", len(x["d2h_bounds"]), "MB up", sum(4 * (e - a) for a, e in x["h2d_runs"]) / 1e6,
      "MB down", sum(4 * (e - a) for a, e in x["d2h_runs"]) / 1e6, "staging pinned", st.is_pinned())


def native(which, direction, threads):
    bounds = x[which + "_bounds"]
    nbytes = np.array([4 * b.n for b in bounds], dtype=np.uint64)
    offs = np.array([4 * (b.off - lo) for b in bounds], dtype=np.uint64)
    ptrs = np.empty(len(bounds), dtype=np.uint64)
    t0 = T()
    for i, b in enumerate(bounds):
        t = b.tensor
        assert t.dtype == torch.float32 and t.numel() == b.n and t.is_contiguous() and not t.is_cuda
        ptrs[i] = t.data_ptr()
    t1 = T()
    rc = sess.lib.dfq_host_copy_segments(C.c_void_p(st.data_ptr()), _lib.table_ptr(ptrs), _lib.table_ptr(nbytes), _lib.table_ptr(offs),
                                         len(bounds), direction, threads)
    assert rc == 0
    return (t1 - t0) * 1e3, (T() - t1) * 1e3


with torch.no_grad():
    for rep in range(3):
        restore(); sync()
        t0 = T(); torch._foreach_copy_(x["h2d_dst"], [b.tensor.detach().reshape(-1) for b in x["h2d_bounds"]]); t1 = T()
        print(f"rep {rep} gather foreach {1e3 * (t1 - t0):.3f} ms")
        for thr in (1, 2, 4, 8, 0):
            a, b = native("h2d", 0, thr)
            print(f"    gather native threads={thr}: pointer loop {a:.3f} ms, copy {b:.3f} ms")
        t0 = T(); torch._foreach_copy_([b.tensor.detach().view(-1) for b in x["d2h_bounds"]], list(x["d2h_src"])); t1 = T()
        print(f"rep {rep} scatter foreach {1e3 * (t1 - t0):.3f} ms")
        for thr in (1, 2, 4, 8, 0):
            a, b = native("d2h", 1, thr)
            print(f"    scatter native threads={thr}: pointer loop {a:.3f} ms, copy {b:.3f} ms")
    for rep in range(4):
        restore(); sync()
        t = [T()]
        cal.upload(); t.append(T()); sync(); t.append(T())
        cal.run_device(equalize=True, correction=True); t.append(T()); sync(); t.append(T())
        cal.download(); sync(); t.append(T())
        names = ["upload()", "h2d wait", "run_device returns", "device done", "download()"]
        print(f"rep {rep}: " + "  ".join(f"{n} {1e3 * (b - a):.3f}" for n, a, b in zip(names, t, t[1:])))
