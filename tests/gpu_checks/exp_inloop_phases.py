"""Where does a step of the unmodified reference training loop spend its time with either bridge? Wall-clock and
CUDA-event time of the bridge forward, the whole backward, and the optimizer part, for this repository's BridgeLite
and the reference's, inside the same FullModel (random-init frozen models, B8 L128). Also: eager fwd+bwd of the bare
bridge before / after a CUDA-graph capture of the same step (allocator interplay)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn as nn

import inloop_harness as H
from vlm_bridge_b200 import BridgeLite, GraphedBridgeStep

layers = int(os.environ.get("GEMMA_LAYERS", "26"))
with H.quiet():
    model = H.build_full_model(BridgeLite, gemma_layers=layers, dino_layers=24, bridge_dropout=0.1)
ours = model.bridge_module
ref = H.reference_bridge_cls()(vision_dim=1024, language_dim=2304, num_heads_cross=8, dropout=0.1).to("cuda")
ref.load_state_dict(ours.state_dict())
batches = H.make_batches(6, 8, 128)


def loop(bridge, tag):
    model.bridge_module = bridge
    model.train()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5, weight_decay=0.01)
    scaler = torch.amp.GradScaler("cuda")
    rec = {"fwd_bridge_host": [], "fwd_bridge_gpu": [], "fwd_total": [], "bwd": [], "unscale_norm_clip": [], "opt": [], "step": []}
    orig_forward = bridge.forward
    ev = []

    def timed_forward(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = orig_forward(*a, **k)
        e1.record()
        rec["fwd_bridge_host"].append((time.perf_counter() - t0) * 1e3)
        ev.append((e0, e1))
        return out

    bridge.forward = timed_forward
    for i, b in enumerate(batches):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        images, ids, mask = b["images"].cuda(), b["input_ids"].cuda(), b["attention_mask"].cuda()
        labels = ids.clone(); labels[:, :-1] = ids[:, 1:]; labels[:, -1] = -100
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(images, ids, mask)["logits"]
            loss = nn.CrossEntropyLoss(ignore_index=-100)(logits.view(-1, logits.size(-1)), labels.view(-1))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        scaler.scale(loss).backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        scaler.unscale_(opt)
        tot = 0.0
        for p in model.parameters():
            if p.requires_grad and p.grad is not None:
                tot += p.grad.data.norm(2).item() ** 2
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.3)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        scaler.step(opt); scaler.update()
        torch.cuda.synchronize(); t4 = time.perf_counter()
        if i >= 2:
            rec["fwd_total"].append((t1 - t0) * 1e3); rec["bwd"].append((t2 - t1) * 1e3)
            rec["unscale_norm_clip"].append((t3 - t2) * 1e3); rec["opt"].append((t4 - t3) * 1e3); rec["step"].append((t4 - t0) * 1e3)
    rec["fwd_bridge_gpu"] = [a.elapsed_time(b) for a, b in ev[2:]]
    rec["fwd_bridge_host"] = rec["fwd_bridge_host"][2:]
    bridge.forward = orig_forward
    print(tag, {k: round(sum(v) / max(1, len(v)), 3) for k, v in rec.items()}, flush=True)


loop(ours, "b200_bridge     ")
loop(ref, "reference_bridge")
loop(ours, "b200_bridge     ")
del model, ref
torch.cuda.empty_cache()

# ---- bare bridge: eager before / after a graph capture ----
g = torch.Generator().manual_seed(1)
v = torch.randn(8, 257, 1024, generator=g).cuda()
t = torch.randn(8, 128, 2304, generator=g).cuda()
params = list(ours.parameters())
ours.train()


def step():
    ours._w16_key = None
    for p in params:
        p.grad = None
    ours(v, t).float().square().mean().backward()


def timed(tag, n=20):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(tag, f"host enqueue {host:.3f} ms/step, wall {(time.perf_counter() - t0) / n * 1e3:.3f} ms/step", flush=True)


timed("eager before graph capture:")
gs = GraphedBridgeStep(ours, lambda y: y.float().square().mean(), v, t)
timed("eager after graph capture: ")
del gs
torch.cuda.empty_cache()
timed("eager after graph deleted: ")
