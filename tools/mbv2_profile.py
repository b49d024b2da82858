"""Per-launch timing of the MobileNetV2 calibration (fold / equalize / correct) with the model resident on the GPU.
usage: [DFQ_TRACE=1] python tools/mbv2_profile.py [topology_name]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn as nn
from dfq_b200 import workload
from dfq_b200.calibrate import GraphCalibration

name = sys.argv[1] if len(sys.argv) > 1 else "mobilenetv2"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
topo = workload.load_topology(os.path.join(root, "tests", "golden", f"topology_{name}.json"))
graph, bottoms, modules = workload.build_graph(topo, seed=0)
cal = GraphCalibration(graph, bottoms, [nn.Conv2d, nn.Linear], device=torch.device("cuda:0"))
cal.upload()
pristine = cal.sess.arena.clone()
sess = cal.sess
print("relations", len(cal.relations), "cle steps", len(cal._cle_plan["step_ptr"]) - 1 if "step_ptr" in cal._cle_plan else "?",
      "bc items", len(cal._bc_items), "levels", 1 + max(i["level"] for i in cal._bc_items))
def ev():
    return torch.cuda.Event(enable_timing=True)
for i in range(4):
    sess.arena.copy_(pristine)
    e = [ev() for _ in range(4)]
    e[0].record(); sess.run_bn_fold(cal._fold_plan)
    e[1].record(); res = sess.run_cle_plan(cal._cle_plan, (1e-8, 1e8), 2e-7, 20, False, 0)
    e[2].record(); sess.run_bias_correct_plan(cal._bc_plan, 8)
    e[3].record(); torch.cuda.synchronize()
    print(f"run {i}: fold {e[0].elapsed_time(e[1]):.3f} ms  equalize {e[1].elapsed_time(e[2]):.3f} ms ({res.n_sweeps} sweeps)  "
          f"correct {e[2].elapsed_time(e[3]):.3f} ms")
