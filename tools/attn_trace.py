"""Timeline of one attention CTA (SM clock stamps of the softmax warps and the MMA thread per KV tile).

    python tools/attn_trace.py build      # here (nvcc, no GPU): fantasy-world_b200/fwb200/libfwb200_trace.so
    python tools/attn_trace.py run        # under gpurun: writes gpurun_out/attn_trace.txt

The instrumented kernel is only in the *_trace.so variant (-DFWB_ATTN_TRACE); libfwb200.so never contains it."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "fantasy-world_b200"
LIB = PKG / "fwb200" / "libfwb200_trace.so"


def build():
    srcs = [PKG / "csrc" / n for n in ("fwb_attn.cu", "fwb_host.cu")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DFWB_ATTN_TRACE", "-shared", "-o", str(LIB), *map(str, srcs)]
    subprocess.run(cmd, check=True)
    print(LIB)


def run():
    import torch

    class T4(C.Structure):
        _fields_ = [("ptr", C.c_void_p), ("sb", C.c_int64), ("sl", C.c_int64), ("sh", C.c_int64)]

    lib = C.CDLL(str(LIB))
    lib.fwb_attn_fwd.argtypes = [C.POINTER(T4)] * 4 + [C.c_int] * 5 + [C.c_float, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.fwb_attn_trace_read.argtypes = [C.c_void_p, C.c_int]
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lib.fwb_attn_set_variant(variant)
    if len(sys.argv) > 3:
        lib.fwb_attn_set_exp2_poly(int(sys.argv[3]))     # pairs of 8 on the exp2 polynomial
    lib.fwb_last_error.restype = C.c_char_p
    out_lines = []
    for (B, H, L, D) in [(1, 40, 32760, 128), (1, 16, 32865, 64)]:
        q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for _ in range(3))
        o = torch.empty_like(q)

        def t4(t):
            return T4(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
        tq, tk, tv, to = t4(q), t4(k), t4(v), t4(o)
        lib.fwb_attn_trace_read(None, 700)          # CTA 700: launched in a later wave, steady state
        for _ in range(2):
            rc = lib.fwb_attn_fwd(C.byref(tq), C.byref(tk), C.byref(tv), C.byref(to), B, H, L, L, D, 1.0 / D ** 0.5, 0, None, 0,
                                  torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.fwb_last_error()
        torch.cuda.synchronize()
        buf = (C.c_longlong * (9 * 64 * 8))()
        lib.fwb_attn_trace_read(buf, -1)
        tr = torch.tensor(list(buf), dtype=torch.int64).view(9, 64, 8)
        t0 = int(tr[0, 8, 0])
        out_lines.append(f"# B{B} H{H} L{L} D{D}: clocks relative to warp 0's wait for KV tile 8")
        out_lines.append("# softmax warp w (tile w//4): wait_begin S_ready ld_done max_done exp_done arrive_done | MMA thread (v1): "
                         "t0: wait_P0 got_P0 issued0 | t1: wait_P1 got_P1 issued1; (attn2: QK / PV issue stamps of the two MMA warps)")
        out_lines.append("# stamps are asm-volatile clock reads: they order memory operations, but ptxas may move register-only math "
                         "across them, so the split between the max / exp columns is approximate; periods are exact")
        for j in range(8, 40):
            row = [f"j={j:2d}"]
            for w in (0, 4):
                row.append(f"w{w}: " + " ".join(f"{int(tr[w, j, e]) - t0:6d}" for e in range(6)))
            row.append("mma: " + " ".join(f"{int(tr[8, j, e]) - t0:6d}" for e in range(6)))
            out_lines.append(" | ".join(row))
        # averages over j = 8..56
        js = slice(8, 56)
        per = (tr[0, 56, 1] - tr[0, 8, 1]).item() / 48
        out_lines.append(f"period per KV tile (warp 0 S_ready to S_ready): {per:.0f} clk")
        for w in range(8):
            d = tr[w, js]
            out_lines.append(f"warp {w}: wait {float((d[:, 1] - d[:, 0]).float().mean()):.0f}  ld {float((d[:, 2] - d[:, 1]).float().mean()):.0f}  "
                             f"max {float((d[:, 3] - d[:, 2]).float().mean()):.0f}  exp {float((d[:, 4] - d[:, 3]).float().mean()):.0f}  "
                             f"store+arrive {float((d[:, 5] - d[:, 4]).float().mean()):.0f}")
        d = tr[8, js]
        out_lines.append(f"mma: wait_P0 {float((d[:, 1] - d[:, 0]).float().mean()):.0f} issue0 {float((d[:, 2] - d[:, 1]).float().mean()):.0f} "
                         f"wait_P1 {float((d[:, 4] - d[:, 3]).float().mean()):.0f} issue1 {float((d[:, 5] - d[:, 4]).float().mean()):.0f}")
        out_lines.append(f"phase: tile1 S_ready minus tile0 S_ready = {float((tr[4, js, 1] - tr[0, js, 1]).float().mean()):.0f} clk")
    text = "\n".join(out_lines)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"attn_trace_v{variant}{'_' + sys.argv[3] if len(sys.argv) > 3 else ''}.txt").write_text(text + "\n")
    print(text)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
