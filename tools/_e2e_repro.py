"""Development helper (GPU box): the e2e arm of bench.py alone, with knobs, to chase a pipeline-only failure.
usage: python tools/_e2e_repro.py <n_chunks> <n_slots> <sync_each 0|1>"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dfq_b200.workload import HostStackCalibrator

n_chunks, n_slots, sync_each = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
hc = HostStackCalibrator(dev, 16, 512, 3, n_slots=n_slots)
n_state = hc.chunk_floats * n_chunks
host_in = torch.empty(n_state, dtype=torch.float32, pin_memory=True)
host_out = torch.empty(n_state, dtype=torch.float32, pin_memory=True)
for st in hc.slots:
    st.generate()
for i in range(n_chunks):
    host_in[i * hc.chunk_floats:(i + 1) * hc.chunk_floats].copy_(hc.slots[i % len(hc.slots)].state())
torch.cuda.synchronize()
if sync_each:
    orig = type(hc.slots[0]).run
    def run(self, *a, **k):
        r = orig(self, *a, **k); torch.cuda.synchronize(); return r
    type(hc.slots[0]).run = run
for rep in range(3):
    t0 = time.time()
    hc.run(host_in, host_out)
    torch.cuda.synchronize()
    print("rep", rep, "ok %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
print("finite", bool(torch.isfinite(host_out).all()))
