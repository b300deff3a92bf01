"""Generate the golden fixtures that pin oracle/bridge_oracle.py to the reference implementation.

Run in the build container only (needs /root/reference, which is absent on the GPU box):
    python tests/golden/make_golden.py
It imports the UNMODIFIED reference `bridge_module.py` by file path and writes
  tests/golden/tiny_bridge.npz      weights, inputs, output, d_text and all 52 parameter gradients
                                    of a tiny-dimension BridgeLite (fp32 CPU, eval mode)
  tests/golden/full_fingerprint.json  scalar fingerprints of the full-dimension module (2304/1024,
                                    2 blocks, 8/18 heads) at B=2, Nv=257, L=64 (config C1): output
                                    statistics, samples, loss, per-parameter gradient norms, plus the
                                    reference's own bf16-autocast deviation from fp32 (tolerance
                                    calibration for the GPU parity tests).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/vlm_bridge/model_architecture/bridge_module.py"


def load_reference():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_bridge_module", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def tiny(ref):
    cfg = dict(vision_dim=32, language_dim=64, num_blocks=2, num_heads_cross=1, num_heads_self=1, dropout=0.0)
    torch.manual_seed(11)
    m = ref.BridgeLite(**cfg).eval()
    # make biases / LayerNorm affine non-trivial so that every gradient path is exercised
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif "ln_" in n and n.endswith("weight"):
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
    g = torch.Generator().manual_seed(13)
    vision = torch.randn(2, 9, 32, generator=g)
    text = torch.randn(2, 5, 64, generator=g).requires_grad_()
    d_out = torch.randn(2, 5, 64, generator=g)
    y = m(vision, text)
    y.backward(d_out)
    out = {"cfg_json": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)}
    for n, p in m.named_parameters():
        out["param/" + n] = p.detach().numpy()
        out["grad/" + n] = p.grad.numpy()
    out.update(vision=vision.numpy(), text=text.detach().numpy(), d_out=d_out.numpy(), y=y.detach().numpy(),
               d_text=text.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "tiny_bridge.npz"), **out)
    print("tiny_bridge.npz written:", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


def full(ref):
    torch.manual_seed(0)
    m = ref.BridgeLite(dropout=0.0).eval()
    g = torch.Generator().manual_seed(1234)
    vision = torch.randn(2, 257, 1024, generator=g)
    text = torch.randn(2, 64, 2304, generator=g).requires_grad_()
    y = m(vision, text)
    loss = y.square().mean()
    loss.backward()
    fp = {
        "torch_version": torch.__version__,
        "config": "C1: B=2 Nv=257 L=64, vision 1024, language 2304, 2 blocks, heads 8/18, eval, fp32 CPU",
        "weights_seed": 0, "inputs_seed": 1234,
        "param_numel": sum(p.numel() for p in m.parameters()),
        "param_abs_sum": {n: float(p.detach().abs().sum()) for n, p in m.named_parameters() if n.endswith("weight")
                          and "ln_" not in n},
        "y_mean": float(y.mean()), "y_std": float(y.std()),
        "y_samples": {"[0,0,:8]": y[0, 0, :8].tolist(), "[1,63,-8:]": y[1, 63, -8:].tolist(),
                      "[1,17,1000:1004]": y[1, 17, 1000:1004].tolist()},
        "loss": float(loss),
        "d_text_norm": float(text.grad.norm()),
        "grad_norms": {n: float(p.grad.norm()) for n, p in m.named_parameters()},
    }
    # the reference's own bf16-autocast deviation from its fp32 result (tolerance calibration)
    m.zero_grad()
    text2 = text.detach().clone().requires_grad_()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        y16 = m(vision, text2)
        loss16 = y16.float().square().mean()
    loss16.backward()
    fp["ref_bf16_autocast_vs_fp32"] = {
        "y_max_abs_over_max_ref": float((y16.float() - y).abs().max() / y.abs().max()),
        "y_max_abs_over_std": float((y16.float() - y).abs().max() / y.std()),
        "loss_abs": abs(float(loss16) - float(loss)),
        "d_text_rel_fro": float((text2.grad - text.grad).norm() / text.grad.norm()),
        "grad_rel_fro": {n: float((p.grad - 0).norm()) for n, p in list(m.named_parameters())[:0]},
    }
    # decode quirk fingerprint: output at earlier positions changes when a token is appended (Fact 2)
    with torch.no_grad():
        y_short = m(vision, text[:, :8].detach())
        y_long = m(vision, text[:, :9].detach())
    fp["noncausal_prefix_change_max_abs"] = float((y_long[:, :8] - y_short).abs().max())
    with open(os.path.join(HERE, "full_fingerprint.json"), "w") as f:
        json.dump(fp, f, indent=1)
    print("full_fingerprint.json written; loss", fp["loss"], "y_std", fp["y_std"],
          "bf16 dev", fp["ref_bf16_autocast_vs_fp32"])


if __name__ == "__main__":
    ref = load_reference()
    tiny(ref)
    full(ref)
