"""State-dict specifications (key -> shape) of the hot-path networks.

These are the reference's checkpoint layouts, which are part of the drop-in boundary
(SURVEY.md §8b): `sample.py:56-70` does `Net(use_cf).load_state_dict(torch.load(path))`.

  denoisers : /root/reference/network.py:1066-1105 (SurfPosNet), :1129-1174 (SurfZNet),
              :1203-1255 (EdgePosNet), :1289-1355 (EdgeZNet); encoder = nn.TransformerEncoder
              (12 pre-norm layers d=768, 12 heads, FFN 1024) + final LayerNorm.
"""
from __future__ import annotations

from typing import List, Tuple

D = 768          # embed_dim, network.py:1073
H = 12           # nhead,     network.py:1076
DH = 64
FF = 1024        # dim_feedforward, network.py:1077
NLAYER = 12      # network.py:1078
NCLASS = 11      # Embedder(11, 768), network.py:1102

# kind -> (ordered embed-MLP names with input widths, output width)
NET_KINDS = {
    "surfpos": ([("p_embed", 6)], 6),
    "surfz": ([("z_embed", 48), ("p_embed", 6)], 48),
    "edgepos": ([("surfz_embed", 48), ("surfp_embed", 6), ("edgep_embed", 6)], 6),
    "edgez": ([("surfz_embed", 48), ("edgez_embed", 12), ("surfp_embed", 6),
               ("edgep_embed", 6), ("vertp_fc", 6)], 18),
}

Spec = List[Tuple[str, Tuple[int, ...]]]


def _mlp(name: str, d_in: int, d_out: int) -> Spec:
    # nn.Sequential(Linear(d_in,768), LayerNorm(768), SiLU, Linear(768,d_out)) -> indices 0,1,3
    return [
        (f"{name}.0.weight", (D, d_in)), (f"{name}.0.bias", (D,)),
        (f"{name}.1.weight", (D,)), (f"{name}.1.bias", (D,)),
        (f"{name}.3.weight", (d_out, D)), (f"{name}.3.bias", (d_out,)),
    ]


def encoder_spec(prefix: str = "net") -> Spec:
    out: Spec = []
    for i in range(NLAYER):
        p = f"{prefix}.layers.{i}"
        out += [
            (f"{p}.self_attn.in_proj_weight", (3 * D, D)), (f"{p}.self_attn.in_proj_bias", (3 * D,)),
            (f"{p}.self_attn.out_proj.weight", (D, D)), (f"{p}.self_attn.out_proj.bias", (D,)),
            (f"{p}.linear1.weight", (FF, D)), (f"{p}.linear1.bias", (FF,)),
            (f"{p}.linear2.weight", (D, FF)), (f"{p}.linear2.bias", (D,)),
            (f"{p}.norm1.weight", (D,)), (f"{p}.norm1.bias", (D,)),
            (f"{p}.norm2.weight", (D,)), (f"{p}.norm2.bias", (D,)),
        ]
    out += [(f"{prefix}.norm.weight", (D,)), (f"{prefix}.norm.bias", (D,))]
    return out


def denoiser_spec(kind: str, use_cf: bool) -> Spec:
    embeds, d_out = NET_KINDS[kind]
    out = encoder_spec("net")
    for name, d_in in embeds:
        out += _mlp(name, d_in, D)
    out += _mlp("time_embed", D, D)
    out += _mlp("fc_out", D, d_out)
    if use_cf:
        out += [("class_embed.embed.weight", (NCLASS, D))]
    return out
