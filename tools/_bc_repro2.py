"""Development helper (GPU box): fold / equalize / correct of a DeviceStack step by step with a sync after each call.
usage: python tools/_bc_repro2.py <n_blocks> <hints 0|1> [channels]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dfq_b200.engine import Session
from dfq_b200.workload import DeviceStack
nb, hints = int(sys.argv[1]), int(sys.argv[2])
C = int(sys.argv[3]) if len(sys.argv) > 3 else 512
sess = Session(torch.device("cuda", 0))
st = DeviceStack(sess, nb, C, 3)
st.generate()
torch.cuda.synchronize()
def step(name, fn):
    t0 = time.time()
    r = fn()
    torch.cuda.synchronize()
    print("%s ok %.2f ms" % (name, (time.time() - t0) * 1e3), flush=True)
    return r
for rep in range(2):
    step("fold", lambda: sess.run_bn_fold(st.fold_plan))
    res = step("cle", lambda: sess.run_cle_plan(st.cle_plan, cols_ready=st.fold_plan["scanned"]))
    h = sess.cle_col_hints(st.cle_plan, res) if hints else None
    step("bc(hints=%d)" % hints, lambda: sess.run_bias_correct_plan(st.bc_plan, 8, col_hints=h))
print("done", bool(torch.isfinite(st.state()).all()))
