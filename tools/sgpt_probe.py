"""Tensor-pipe side probe: bench.sparsegpt_roofline() alone (Hessian shapes, OBS phases)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
r = bench.sparsegpt_roofline(torch.device("cuda", 0))
for h in r["hessian"]:
    print(f"hessian {h['shape']:50s} {h['ms']*1e3:8.1f} us  alg {h['tflops_algorithmic']:7.1f} TF/s  executed {h['tflops_executed']:7.1f} ({h['tensor_pipe_frac_of_sustained']:.2f} of sustained)")
for o in r["obs"]:
    print(f"obs {o['shape']:28s} total {o['total_ms']:.3f} ms = thr {o['threshold_ms']:.3f} + sweep {o['sweep_ms']:.3f} + trailing {o['trailing_ms']:.3f} ({o['trailing_tflops_algorithmic']:.1f} TF/s alg); prologue {o['prologue_cusolver_ms']:.1f} ms")
json.dump(r, open("gpurun_out/sgpt_probe.json", "w"))
