#!/bin/bash
# GPU call J (1 GPU): persistent row kernels — correctness, CTAs-per-SM sweep, in-step breakdown
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "ln_modulate or rmsnorm or ln64 or cfg_euler or rowwise" 2>&1 | tail -5)
python - <<'PY'
import sys; sys.path.insert(0, "fantasy-world_b200")
import torch, fwb200
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
x = torch.randn(32760, 5120, device="cuda").to(torch.bfloat16)
xs = [x.clone() for _ in range(4)]            # rotate buffers: 4 x 335 MB > L2
mul = torch.randn(5120, device="cuda"); add = torch.randn(5120, device="cuda")
out = torch.empty_like(x)
cs = torch.randn(32760, 64, 2, device="cuda")
w = torch.randn(5120, device="cuda")
xf = [torch.randn(32865, 1024, device="cuda") for _ in range(8)]
wf = torch.randn(1024, device="cuda"); bf = torch.randn(1024, device="cuda")
outf = torch.empty(32865, 1024, device="cuda", dtype=torch.bfloat16)
for n in (0, 4, 6, 8, 12):
    fwb200.lib.fwb_rowwise_set_ctas_per_sm(n)
    i = [0]
    def ln():
        i[0] += 1; fwb200.ln_modulate(xs[i[0] % 4], eps=1e-6, mul=mul, add=add, out=out)
    def rr():
        i[0] += 1; fwb200.rmsnorm_rope_(xs[i[0] % 4], w=w, eps=1e-6, cos_sin=cs, head_dim=128)
    def lf():
        i[0] += 1; fwb200.ln_modulate(xf[i[0] % 8], eps=1e-5, w=wf, b=bf, out=outf)
    t1, t2, t3 = timeit(ln), timeit(rr), timeit(lf)
    gb1 = 32760 * 5120 * 4 / 1e9; gb3 = 32865 * 1024 * 6 / 1e9
    print(f"ctas/SM {n:2d}: ln_modulate[32760,5120] {t1*1e3:6.1f} us {gb1/t1:6.2f} TB/s | rmsnorm_rope {t2*1e3:6.1f} us {(gb1 + 32760*64*8/1e9)/t2:6.2f} TB/s | ln fp32[32865,1024] {t3*1e3:6.1f} us {gb3/t3:6.2f} TB/s", flush=True)
fwb200.lib.fwb_rowwise_set_ctas_per_sm(0)
PY
