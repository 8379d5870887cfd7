"""In-step A/B of a tuning setter: ONE process, ONE model, alternating segments of full denoise steps with the setter at each value
(the board is power-capped, so isolated kernel timings do not predict the step; box-to-box clocks differ, so the comparison has to
happen inside one run).

    python tools/gpu_step_ab.py fwb_attn_set_multicast 0 1 [--rounds 3] [--steps 3]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
import torch
import fwb200
from fwb200.synth import build_fusion_model, synth_inputs

ap = argparse.ArgumentParser()
ap.add_argument("setter")
ap.add_argument("values", type=int, nargs="+")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = build_fusion_model(num_dit_layers=40, start_index=16, device=dev, seed=0, heads=False)
model.pipe.device = dev
inp = synth_inputs(21, 30, 52, device=dev, seed=1024, text_len=512)
lens = torch.ones(21, dtype=torch.long, device=dev)
lens[1:] = 4
model.pipe.scheduler.set_timesteps(50)
fn = getattr(fwb200.lib, a.setter)


def steps(n, lat, i0):
    for i in range(n):
        lat = model.denoise_step(lat, (i0 + i) % 50, inp["context_pos"], inp["context_neg"], clip_feature=inp["clip_feature"], y=inp["y"],
                                 plucker_fea=inp["plucker_fea"], plucker_context_lens=lens, cfg_scale=5.0)[0]
    return lat


lat = steps(3, inp["latents"].clone(), 0)          # warm-up
torch.cuda.synchronize()
res = {v: [] for v in a.values}
i0 = 3
for r in range(a.rounds):
    for v in a.values:
        fn(v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = steps(a.steps, lat, i0)
        e1.record()
        torch.cuda.synchronize()
        i0 += a.steps
        res[v].append(e0.elapsed_time(e1) / a.steps)
fn(a.values[0])
out = f"{a.setter}: " + "   ".join(f"value {v}: " + "/".join(f"{m:.0f}" for m in res[v]) + f" ms/step (mean {sum(res[v]) / len(res[v]):.1f})" for v in a.values)
print(out)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
with open(ROOT / "gpurun_out" / "r02_step_ab.log", "a") as f:
    f.write(out + "\n")
