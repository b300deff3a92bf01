"""CPU oracle of the bridge hot path -- TEST INFRASTRUCTURE ONLY.

A functional restatement, in plain fp32 torch-on-CPU primitives (matmul, softmax, erf, mean/var), of
the reference `BridgeLite` (src/vlm_bridge/model_architecture/bridge_module.py). It exists so the
CUDA path can be checked on a GPU box where /root/reference is absent. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it; the
product package never does (tests/test_no_oracle_in_product.py enforces that).

Where the arithmetic lives: the reference delegates every op to PyTorch (pinned torch==2.7.1,
uv.lock:1165-1166; this image has 2.11.0): nn.Linear, nn.LayerNorm, F.scaled_dot_product_attention,
nn.GELU (erf), nn.Dropout. Their published definitions are restated here.

Pinning: the reference ships NO golden vectors for this path (its one bridge test asserts a shape,
test_model_architecture.py:142-144). The oracle is therefore pinned against outputs of the
reference module itself, imported unmodified in the build container by
tests/golden/make_golden.py; the resulting fixtures (tests/golden/*.npz, *.json) are committed and
tests/test_oracle_golden.py checks the oracle against them on every run.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor


def _bf16(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
def param_names(num_blocks: int = 2) -> list[str]:
    """The 26*num_blocks state_dict keys in registration order (SURVEY.md Appendix A;
    bridge_module.py:275-298)."""
    names = []
    for i in range(num_blocks):
        pre = f"bridge_blocks.{i}."
        for lin in ("w_q", "w_k", "w_v", "w_o"):
            names += [pre + f"cross_attention.{lin}.weight", pre + f"cross_attention.{lin}.bias"]
        names += [pre + "ln_cross.weight", pre + "ln_cross.bias"]
        for lin in ("w_q", "w_k", "w_v", "w_o"):
            names += [pre + f"self_attention.{lin}.weight", pre + f"self_attention.{lin}.bias"]
        names += [pre + "ln_self.weight", pre + "ln_self.bias"]
        names += [pre + "ffn.0.weight", pre + "ffn.0.bias", pre + "ffn.3.weight", pre + "ffn.3.bias"]
        names += [pre + "ln_ffn.weight", pre + "ln_ffn.bias"]
    return names


def init_state_dict(seed: int, vision_dim: int = 1024, language_dim: int = 2304, num_blocks: int = 2) -> Dict[str, Tensor]:
    """Weights the reference constructor produces after `torch.manual_seed(seed)`.

    The reference first builds every nn.Linear (default Kaiming-uniform init, which consumes the
    global RNG) in the order cross w_q,w_k,w_v,w_o, self w_q,w_k,w_v,w_o, ffn.0, ffn.3 per block
    (bridge_module.py:65-70,168-173,291-297), then `_init_weights` (:394-404) walks the modules in the
    same order re-drawing weights Xavier-uniform, zeroing biases, resetting LayerNorm to (1, 0).
    The RNG stream is restated draw for draw.
    """
    D, Dv, F = language_dim, vision_dim, 4 * language_dim
    shapes = []
    for _ in range(num_blocks):
        shapes += [(D, D), (D, Dv), (D, Dv), (D, D), (D, D), (D, D), (D, D), (D, D), (F, D), (D, F)]
    torch.manual_seed(seed)
    # pass 1: nn.Linear.reset_parameters = kaiming_uniform_(weight, a=sqrt(5)) then uniform_(bias)
    for out_f, in_f in shapes:
        torch.empty(out_f, in_f).uniform_(-1.0, 1.0)  # one uniform_ draw of weight.numel() values
        torch.empty(out_f).uniform_(-1.0, 1.0)        # one uniform_ draw of bias.numel() values
    # pass 2: xavier_uniform_ in module order
    weights = []
    for out_f, in_f in shapes:
        bound = math.sqrt(6.0 / (in_f + out_f))
        weights.append(torch.empty(out_f, in_f).uniform_(-bound, bound))
    sd: Dict[str, Tensor] = {}
    it = iter(weights)
    for i in range(num_blocks):
        pre = f"bridge_blocks.{i}."
        for grp in ("cross_attention", "self_attention"):
            for lin in ("w_q", "w_k", "w_v", "w_o"):
                w = next(it)
                sd[pre + f"{grp}.{lin}.weight"] = w
                sd[pre + f"{grp}.{lin}.bias"] = torch.zeros(w.shape[0])
        for lin in ("ffn.0", "ffn.3"):
            w = next(it)
            sd[pre + lin + ".weight"] = w
            sd[pre + lin + ".bias"] = torch.zeros(w.shape[0])
        for ln in ("ln_cross", "ln_self", "ln_ffn"):
            sd[pre + ln + ".weight"] = torch.ones(D)
            sd[pre + ln + ".bias"] = torch.zeros(D)
    return {k: sd[k] for k in param_names(num_blocks)}


# ------------------------------------------------------------------------------------------------
# primitive ops, restated
# ------------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance, eps inside the sqrt (bridge_module.py:282)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def linear(x: Tensor, w: Tensor, b: Tensor, emulate_bf16: bool) -> Tensor:
    """y = x W^T + b (nn.Linear). Under autocast operands and result are bf16 (Appendix B)."""
    if emulate_bf16:
        return _bf16(_bf16(x) @ _bf16(w).t() + _bf16(b))
    return x @ w.t() + b


def gelu(x: Tensor) -> Tensor:
    """Exact GELU, 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU() default, bridge_module.py:293)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def attention(q: Tensor, k: Tensor, v: Tensor, heads: int, emulate_bf16: bool) -> Tensor:
    """softmax(Q K^T / sqrt(d_k)) V per head, no mask, non-causal (bridge_module.py:103-115,132-139)."""
    B, Lq, D = q.shape
    Lk = k.shape[1]
    d = D // heads
    qh = q.view(B, Lq, heads, d).transpose(1, 2)
    kh = k.view(B, Lk, heads, d).transpose(1, 2)
    vh = v.view(B, Lk, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    o = p @ vh
    o = o.transpose(1, 2).reshape(B, Lq, D)
    return _bf16(o) if emulate_bf16 else o


# ------------------------------------------------------------------------------------------------
# the bridge
# ------------------------------------------------------------------------------------------------
def block_forward(sd: Dict[str, Tensor], i: int, x: Tensor, vision: Tensor, heads_cross: int, heads_self: int,
                  emulate_bf16: bool = False, kv: Optional[tuple] = None) -> Tensor:
    """One BridgeBlock in eval mode (bridge_module.py:300-335)."""
    pre = f"bridge_blocks.{i}."

    def P(n):
        return sd[pre + n]

    # cross-attention sub-layer (:316-323)
    xn = layer_norm(x, P("ln_cross.weight"), P("ln_cross.bias"))
    q = linear(xn, P("cross_attention.w_q.weight"), P("cross_attention.w_q.bias"), emulate_bf16)
    if kv is None:
        k = linear(vision, P("cross_attention.w_k.weight"), P("cross_attention.w_k.bias"), emulate_bf16)
        v = linear(vision, P("cross_attention.w_v.weight"), P("cross_attention.w_v.bias"), emulate_bf16)
    else:
        k, v = kv
    a = attention(q, k, v, heads_cross, emulate_bf16)
    x = x + linear(a, P("cross_attention.w_o.weight"), P("cross_attention.w_o.bias"), emulate_bf16)
    # self-attention sub-layer (:326-328) -- non-causal, unmasked
    xn = layer_norm(x, P("ln_self.weight"), P("ln_self.bias"))
    q = linear(xn, P("self_attention.w_q.weight"), P("self_attention.w_q.bias"), emulate_bf16)
    k = linear(xn, P("self_attention.w_k.weight"), P("self_attention.w_k.bias"), emulate_bf16)
    v = linear(xn, P("self_attention.w_v.weight"), P("self_attention.w_v.bias"), emulate_bf16)
    a = attention(q, k, v, heads_self, emulate_bf16)
    x = x + linear(a, P("self_attention.w_o.weight"), P("self_attention.w_o.bias"), emulate_bf16)
    # FFN sub-layer (:331-333)
    xn = layer_norm(x, P("ln_ffn.weight"), P("ln_ffn.bias"))
    h = gelu(linear(xn, P("ffn.0.weight"), P("ffn.0.bias"), emulate_bf16))
    if emulate_bf16:
        h = _bf16(h)
    x = x + linear(h, P("ffn.3.weight"), P("ffn.3.bias"), emulate_bf16)
    return x


def bridge_forward(sd: Dict[str, Tensor], vision: Tensor, text: Tensor, *, num_blocks: int = 2, heads_cross: int = 8,
                   heads_self: int = 18, emulate_bf16: bool = False, return_blocks: bool = False):
    """BridgeLite.forward in eval mode / dropout 0 (bridge_module.py:406-456): every block sees the
    same raw `vision` features."""
    x = text
    outs = []
    for i in range(num_blocks):
        x = block_forward(sd, i, x, vision, heads_cross, heads_self, emulate_bf16)
        outs.append(x)
    return (x, outs) if return_blocks else x


def vision_kv(sd: Dict[str, Tensor], vision: Tensor, num_blocks: int = 2, emulate_bf16: bool = False):
    """Per-block (K, V) of the image: the only exactly cacheable decode state (SURVEY.md Fact 2)."""
    out = []
    for i in range(num_blocks):
        pre = f"bridge_blocks.{i}.cross_attention."
        out.append((linear(vision, sd[pre + "w_k.weight"], sd[pre + "w_k.bias"], emulate_bf16),
                    linear(vision, sd[pre + "w_v.weight"], sd[pre + "w_v.bias"], emulate_bf16)))
    return out


def bridge_forward_cached(sd, kvs, text, *, num_blocks=2, heads_cross=8, heads_self=18, emulate_bf16=False):
    """Decode-time forward over cached vision K/V; equals bridge_forward on the same image."""
    x = text
    for i in range(num_blocks):
        x = block_forward(sd, i, x, None, heads_cross, heads_self, emulate_bf16, kv=kvs[i])
    return x


def bridge_loss_and_grads(sd: Dict[str, Tensor], vision: Tensor, text: Tensor, *, num_blocks: int = 2,
                          heads_cross: int = 8, heads_self: int = 18, emulate_bf16: bool = False,
                          d_out: Optional[Tensor] = None):
    """Forward + backward through the restated ops (torch autograd over the primitives above).

    Upstream gradient: `d_out` if given, else that of loss = mean(y^2).
    Returns (y, loss, d_text, {name: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    t = text.detach().clone().requires_grad_(True)
    y = bridge_forward(leaves, vision, t, num_blocks=num_blocks, heads_cross=heads_cross, heads_self=heads_self,
                       emulate_bf16=emulate_bf16)
    loss = y.float().square().mean()
    if d_out is None:
        loss.backward()
    else:
        y.backward(d_out)
    return y.detach(), float(loss.detach()), t.grad, {k: v.grad for k, v in leaves.items()}


def greedy_decode_bridge_only(sd, vision, embed: Tensor, head: Tensor, steps: int, bos: int = 2, **kw):
    """Greedy caption loop of FullModel.generate_caption (full_model.py:241-363) with the frozen LM
    replaced by a fixed linear read-out: embed [V, D] plays get_embeddings, `head` [V, D] the LM.
    Every step recomputes the bridge over the whole prefix (self-attention is non-causal, Fact 2).
    Returns the [B, steps+1] token ids."""
    B = vision.shape[0]
    ids = torch.full((B, 1), bos, dtype=torch.long)
    for _ in range(steps):
        x = embed[ids]
        y = bridge_forward(sd, vision, x, **kw)
        logits = y[:, -1, :] @ head.t()
        ids = torch.cat([ids, logits.argmax(-1, keepdim=True)], dim=1)
    return ids
