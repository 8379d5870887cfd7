"""Diagnostic (GPU): why the full-size umT5 test uses the encoder's own initialisation instead of the per-key synthetic weights.
With unit-variance q and k (fwb_synth) and no 1/sqrt(d) in T5's logits the softmaxes are near one-hot; this prints, per block and on
the SAME input, kernels-vs-shim error, and the end-to-end error, for both initialisations."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT / "fantasy-world_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

from _ops_torch_shim import torch_ops                                        # noqa: E402
from FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder import WanTextEncoder, init_weights   # noqa: E402
from fwb_synth import materialize, synth_init                                # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 256384, (1, 512), generator=g).cuda()
    mask = torch.zeros(1, 512, dtype=torch.long, device="cuda")
    mask[:, :37] = 1
    for init in ("synth", "own"):
        with torch.device("meta"):
            m = WanTextEncoder(num_layers=layers)
        m = materialize(m, "cuda", torch.bfloat16).eval()
        with torch.no_grad():
            if init == "synth":
                synth_init(m, seed=0)
            else:
                torch.manual_seed(0)
                m.apply(init_weights)
                torch.nn.init.normal_(m.token_embedding.weight)
            x = m.token_embedding(ids).to(torch.bfloat16)
            per = []
            for blk in m.blocks:
                y = blk(x, mask)
                with torch_ops():
                    yr = blk(x, mask)
                per.append(rel(y[:, :37], yr[:, :37]))
                x = y
            out = m(ids, mask)
            with torch_ops():
                ref = m(ids, mask)
        print(f"init={init} layers={layers} end-to-end {rel(out[:, :37], ref[:, :37]):.3e} per-block(same input) "
              + " ".join(f"{e:.1e}" for e in per), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
