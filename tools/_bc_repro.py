"""Development helper (GPU box): DeviceStack(n_blocks).run() repeated, on the default or a side stream.
usage: python tools/_bc_repro.py <n_blocks> <side_stream 0|1> [reps]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dfq_b200.engine import Session
from dfq_b200.workload import DeviceStack
nb, side = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sess = Session(torch.device("cuda", 0))
st = DeviceStack(sess, nb, 512, 3)
st.generate()
saved = st.state().clone()
torch.cuda.synchronize()
stream = torch.cuda.Stream() if side else torch.cuda.current_stream()
with torch.cuda.stream(stream):
    for r in range(reps):
        st.state().copy_(saved)
        t0 = time.time()
        res = st.run()
        stream.synchronize()
        print("blocks", nb, "side", side, "rep", r, "ok %.2f ms sweeps %d" % ((time.time() - t0) * 1e3, res.n_sweeps), flush=True)
