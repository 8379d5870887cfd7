"""bench.py — denoise-steps/sec of the FantasyWorld hot path (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py --gpus 1 --steps 3 --warmup 3                     # our CUDA path (default workload = configs[1])
    python bench.py --impl reference --steps 1 --warmup 0             # the reference's CPU path (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...  # N > 1: see DESIGN.md §multi-GPU

Workload (configs[1]): Wan2.1-I2V-14B-480P shape, random-init, latents [1,16,21,60,104] (81 frames, 480x832),
16 Preconditioning Blocks + 24 IRG blocks, one step = conditional + unconditional joint_forward + CFG + Euler update
(geometry heads excluded from the per-step metric).  Synthetic inputs, synthetic weights.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same step through the public module API
(FantasyWorldFusionModel.denoise_step) from pinned HOST buffers, H2D of the step's latents and D2H of the result inside
the timed region (conditioning tensors are uploaded once per run inside the timed region and amortised over the steps).
`roofline`: the dominant kernel (DiT self-attention, fwb_attn_fwd D=128) timed per launch with CUDA events during the
timed steps.  `cpu_baseline`: oracle port on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "fantasy-world_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

STEP_FLOP_C2 = 4.267e15          # algorithmic FLOP per denoise step at C2 (SURVEY Appendix C / BASELINE.md §2)


def forward_flops(f, h, w, n_pcb, n_irg, text_len=512):
    """Algorithmic FLOPs of one joint_forward (attention 4*H*Lq*Lk*D, GEMM 2*M*N*K), same accounting as SURVEY App. C."""
    L, P = f * h * w, 5 + h * w
    N = f * P
    C, Fd = 5120, 13824
    dit = 2 * L * C * C * 4 + 4 * 40 * L * L * 128                        # qkvo + self-attn
    dit += 2 * L * C * C * 2 + 4 * 40 * L * (text_len + 257) * 128         # cross q,o + attn (K/V of the context are hoisted)
    dit += 2 * L * C * Fd * 2                                              # FFN
    cam = 2 * L * (C * 1024 + 1024 * 2048 + 2048 * 409 + 409 * C)          # camera AdaLN MLPs (group1 hoisted)
    vg_lin = 2 * N * 1024 * (3072 + 1024 + 4096 + 4096)
    frame = vg_lin + 4 * 16 * f * P * P * 64
    glob = vg_lin + 4 * 16 * N * N * 64
    adapter = 2 * L * C * 2304 + 2 * N * 1024 * 2304 + 2 * L * 1152 * C + 2 * N * 1152 * 1024 + 2 * 4 * 12 * L * N * 96
    total = n_pcb * (dit + cam) + n_irg * (frame + dit + glob + adapter) + min(n_irg, 9) * cam
    total += 2 * L * C * 1024                                              # projection head
    return float(total)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu_index, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's algorithm on the host cores (bounded sample)
# ----------------------------------------------------------------------------------------------------------------------
def usable_cores():
    """Host threads this process can really use: the affinity mask, capped by the cgroup CPU quota when there is one
    (os.cpu_count() reports the machine, and oversubscribing a quota makes the CPU arm look worse than it is)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = Path(path).read_text().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


_CPU_SD = None


def cpu_sample(f=1, h=30, w=52, text_len=512, reps=1):
    """One PCB DiT block + one VGGT frame block + one IRG block at f,h,w (full 14B widths), fp32, all host threads.
    Returns (seconds, algorithmic FLOPs, description)."""
    import torch
    from fwb200.synth import synth_tensor
    from oracle import fw_oracle as O
    torch.set_num_threads(usable_cores())
    schema = json.loads((ROOT / "tests" / "golden" / "schema_reduced.json").read_text())
    keys = [k for k in schema if k.startswith(("pipe.dit.blocks.0.", "vggt.aggregator.frame_blocks.0.", "IRGBlock.0.",
                                                "vggt.aggregator.camera_token", "vggt.aggregator.register_token"))]
    global _CPU_SD
    if _CPU_SD is None:                                  # weights are generated once per process
        _CPU_SD = {k: synth_tensor(k, schema[k], 0, "cpu") for k in keys}
    sd = _CPU_SD
    g = torch.Generator().manual_seed(1024)
    L, P = f * h * w, 5 + h * w
    x = torch.randn(1, L, 5120, generator=g)
    tok = torch.randn(f, P, 1024, generator=g)
    ctx = torch.randn(1, 257 + text_len, 5120, generator=g)
    t_mod = torch.randn(1, 6, 5120, generator=g) * 0.1
    e0 = torch.randn(1, 6, 1024, generator=g) * 0.1
    plucker = torch.randn(1, L, 2048, generator=g)
    tab, tab_d, tab_a = O.rope_table_3d(128, f, h, w), O.rope_table_3d(96, f, h, w), O.rope_table_3d_with_extra(96, f, h, w, 5)
    _, pos = O.aggregator_input(sd, "vggt.aggregator", torch.zeros(1, f, h, w, 1024))
    O.USE_TORCH_SDPA = True      # time what the reference runs on CPU (F.scaled_dot_product_attention), not the explicit restatement
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(reps):
            x1 = O.dit_block(sd, "pipe.dit.blocks.0", x, ctx, t_mod, tab, plucker)
            tk = O.vggt_block(sd, "vggt.aggregator.frame_blocks.0", tok, pos, e0)
            O.irg_block(sd, "IRGBlock.0", x1, tk, ctx, t_mod, tab, tab_d, tab_a, pos, e0, plucker)
    dt = (time.perf_counter() - t0) / reps
    O.USE_TORCH_SDPA = False
    # the oracle recomputes the context K/V inside every block (as the reference does): count them
    extra_kv = 2 * 2 * (text_len + 257) * 5120 * 5120 * 2 + 2 * 2 * L * 2048 * 2048
    flops = forward_flops(f, h, w, 1, 1, text_len) - 2 * L * 5120 * 1024 + extra_kv
    return dt, flops, f"1 PCB DiT block + 1 VGGT frame block + 1 IRG block at f,h,w={f},{h},{w} (L={L}), fp32 oracle port"


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    times, flops, desc = [], 0.0, ""
    # every step is one bounded sample; its token count is chosen once so that warmup + steps end within ~4 minutes on this host
    # (the metric is FLOP-rate based, so a smaller sample measures the same thing)
    h, w = 30, 52
    t_probe = time.perf_counter()
    dt0, _, _ = cpu_sample(h=h, w=w)
    n = args.warmup + args.steps
    while dt0 * n > 240.0 and h > 4:
        h, w, dt0 = h // 2, w // 2, dt0 / 4.0
    probe_s = time.perf_counter() - t_probe
    for i in range(n):
        dt, flops, desc = cpu_sample(h=h, w=w, reps=1)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    desc += f" (probe {probe_s:.1f} s)"
    rate = flops / dt                                   # FLOP/s of the reference algorithm on this host
    steps_per_s = rate / STEP_FLOP_C2
    line = {"metric": "denoise_steps_per_sec", "value": steps_per_s, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / steps_per_s, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "Wan2.1-I2V-14B-480P shape, latents 1x16x21x60x104 (81 frames), 16 PCB + 24 IRG, CFG 2 forwards/step",
                       "note": "each timed step is a bounded sample; steps/s = measured FLOP/s / 4.267 PFLOP per step"},
            "cpu_baseline": {"value": steps_per_s, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc,
                             "sample_seconds": dt, "achieved_tflops": rate / 1e12},
            "e2e": {"value": steps_per_s, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------------
# GPU path
# ----------------------------------------------------------------------------------------------------------------------
def dominant_tag(L, world):
    """Profiling-tag prefix of the dominant kernel's launches: the DiT self-attention of this rank's L / world query rows.
    Under sequence parallelism every attention runs as split-KV partials over slices of the keys, and the slices may be ragged
    (4095 rows per rank in 4 slices = 1024, 1024, 1024, 1023), so the match is on the query side only."""
    return f"attn:B1:H40:Lq{L // world}:Lk"


def roofline_from_prof(prof, L, world, step_ms_total, sustained_peak, peak_src):
    """The `roofline` object from the CUDA-event records {tag: (launches, total_ms)} of the timed region (pure function)."""
    tag = dominant_tag(L, world)

    def lk(t):
        return int(t.split(":Lk")[1].split(":")[0])

    # the text / CLIP cross-attentions share the prefix (same q) but have 512 / 257 keys: not the dominant kernel
    dom = {t: v for t, v in prof.items() if t.startswith(tag) and t.endswith(":D128") and lk(t) > 1024}
    if not dom:
        return None
    cnt = sum(c for c, _ in dom.values())
    tot = sum(ms_ for _, ms_ in dom.values())
    fl_total = sum(c * 4.0 * 40 * (L // world) * lk(t) * 128 for t, (c, _) in dom.items())
    ach = fl_total / (tot * 1e-3) / 1e12
    traffic = None
    tp = ROOT / "profiles" / "attn_d128_dram_bytes.json"
    if tp.exists() and world == 1 and L == 32760:        # the ncu capture is of the full-size single-GPU launch
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    return {"kernel": "attn_fwd_kernel<128> (DiT self-attention, fwb_attn_fwd / fwb_attn_fwd_partial)", "bound": "tensor",
            "achieved": ach, "peak": sustained_peak, "unit": "TFLOP/s", "frac": ach / sustained_peak, "traffic": traffic,
            "launches_timed": cnt, "ms_per_launch": tot / cnt, "flop_per_launch": fl_total / cnt,
            "peak_source": peak_src + ", sustained figure (kernel timed inside a long step)",
            "share_of_step": tot / step_ms_total}


def run_ours(args):
    import torch
    import fwb200
    from fwb200.synth import build_fusion_model, synth_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    fwb200.require_device()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    f, h, w = args.frames, args.h, args.w
    n_pcb, n_irg = args.pcb, args.irg
    model = build_fusion_model(num_dit_layers=n_pcb + n_irg, start_index=n_pcb, device=dev, seed=0, heads=False)
    model.pipe.device = dev
    inp = synth_inputs(f, h, w, device=dev, seed=1024, text_len=512)       # one sample, replicated inputs
    if world > 1:
        from fwb200.sp import SPContext
        model.sp = SPContext()                                               # tokens sharded over the ranks
    lens = torch.ones(f, dtype=torch.long, device=dev)
    lens[1:] = 4
    sched = model.pipe.scheduler
    sched.set_timesteps(50)
    n_sched = len(sched.timesteps)

    def one_step(lat, i):
        return model.denoise_step(lat, i % n_sched, inp["context_pos"], inp["context_neg"], clip_feature=inp["clip_feature"],
                                  y=inp["y"], plucker_fea=inp["plucker_fea"], plucker_context_lens=lens, cfg_scale=5.0)[0]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lat = inp["latents"].clone()
    for i in range(args.warmup):
        lat = one_step(lat, i)
    barrier()

    # ---- timed region: device-resident inputs --------------------------------------------------------------------------
    L = f * h * w
    dom_tag = dominant_tag(L, world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fwb200.reset_launch_count()
    fwb200.prof_enable(prefixes=[dom_tag] if not args.breakdown else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        lat = one_step(lat, args.warmup + i)
    e1.record()
    barrier()
    launches = fwb200.launch_count()
    prof = fwb200.prof_disable()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    sampler.stop_flag.set()

    # ---- e2e: host buffers through the public API ------------------------------------------------------------------------
    host = {k: v.cpu().pin_memory() for k, v in inp.items()}
    lat_host = host["latents"].clone().pin_memory()
    out_host = torch.empty_like(lat_host).pin_memory()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    cond = {k: host[k].to(dev, non_blocking=True) for k in ("context_pos", "context_neg", "clip_feature", "y", "plucker_fea")}
    h2d = sum(host[k].numel() * host[k].element_size() for k in cond)
    d2h = 0
    for i in range(args.steps):
        cur = model.denoise_step(lat_host, (args.warmup + i) % n_sched, cond["context_pos"], cond["context_neg"],
                                 clip_feature=cond["clip_feature"], y=cond["y"], plucker_fea=cond["plucker_fea"],
                                 plucker_context_lens=lens, cfg_scale=5.0)[0]
        out_host.copy_(cur, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the host needs the step's result before it can feed the next step
        lat_host.copy_(out_host)
        h2d += lat_host.numel() * lat_host.element_size()
        d2h += out_host.numel() * out_host.element_size()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    if dist is not None:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)

    if rank != 0:
        return
    steps_per_s = args.steps / (ms / 1e3)               # strong scaling: all ranks denoise ONE sample (sequence parallel)
    e2e_sps = args.steps / (ms_e2e / 1e3)
    burst, sustained, peak_src = peaks()
    roof = roofline_from_prof(prof, L, world, ms, sustained, peak_src)
    fwd_fl = forward_flops(f, h, w, n_pcb, n_irg)
    cpu = None
    if not args.no_cpu_baseline:
        dt, cfl, desc = cpu_sample()
        if dt < 6.0:                                     # fast host: take a larger sample (10-30 s of CPU work)
            dt, cfl, desc = cpu_sample(f=min(4, int(12.0 / dt) + 1))
        rate = cfl / dt
        cpu = {"value": rate / (2 * fwd_fl), "unit": "steps/s", "cores": usable_cores(), "kind": "port", "sample": desc,
               "sample_seconds": dt, "achieved_tflops": rate / 1e12}
    full = (f, h, w, n_pcb, n_irg) == (21, 30, 52, 16, 24)
    line = {"metric": "denoise_steps_per_sec", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("Wan2.1-I2V-14B-480P shape (BASELINE configs[1]): " if full else "REDUCED (not the headline config): ") +
                       f"latents 1x16x{f}x{2 * h}x{2 * w}, {n_pcb} PCB + {n_irg} IRG blocks, 2 forwards/step (CFG 5.0), random-init",
                       "tokens_video": L, "tokens_geometry": f * (5 + h * w), "flop_per_step": 2 * fwd_fl,
                       "achieved_tflops_per_gpu": 2 * fwd_fl * args.steps / (ms / 1e3) / 1e12 / world,
                       "parallelism": f"sp{world} (token-sharded sequence parallel, 1 all-gather of packed K|V per attention)" if world > 1 else "single",
                       "sp_gathers_per_step": (model.sp.n_gathers / max(1, args.warmup + 2 * args.steps)) if world > 1 else 0,
                       "sp_gather_bytes_per_step": (model.sp.gather_bytes / max(1, args.warmup + 2 * args.steps)) if world > 1 else 0,
                       "l2": "per-step working set (37 GB weights + >1 GB activations) exceeds the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_sps, "unit": "steps/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "api": "FantasyWorldFusionModel.denoise_step from pinned host latents; conditioning uploaded once per run inside the timed region"},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}
    if args.breakdown:
        agg = sorted(((t, c, ms_) for t, (c, ms_) in prof.items()), key=lambda r: -r[2])
        line["breakdown_ms_per_step"] = [{"tag": t, "launches_per_step": c / args.steps, "ms_per_step": m / args.steps} for t, c, m in agg[:40]]
    emit(line)


_JSON_FD = None


def protect_stdout():
    """Libraries (NCCL's version banner, torch warnings) may write to fd 1; the driver expects ONE JSON line there.  Keep a
    private copy of stdout for the result and point fd 1 at stderr for everything else."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=21)
    ap.add_argument("--h", type=int, default=30)
    ap.add_argument("--w", type=int, default=52)
    ap.add_argument("--pcb", type=int, default=16)
    ap.add_argument("--irg", type=int, default=24)
    ap.add_argument("--breakdown", action="store_true", help="time every fwb200 launch with CUDA events and add a per-kernel table")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
