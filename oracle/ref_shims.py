"""TEST INFRASTRUCTURE (oracle side) -- shims that let the UNMODIFIED reference (/root/reference) import here.

The reference depends on timm==0.9.16, openai/CLIP, einops_exts and two symbols transformers>=5 no longer exports;
none are installed in this image.  `install()` registers minimal stand-ins in sys.modules that restate
the arithmetic of those third-party pieces, then puts /root/reference on sys.path so `models.dreamvla_model` etc.
import unchanged.  Used ONLY by tests/golden/make_golden.py and CPU tests that pin the oracle against the reference
(skipped when /root/reference is absent, e.g. on the GPU box).  Never imported by the product path.

Restated third-party arithmetic (pinned versions from the reference's requirements.txt):
  timm 0.9.16  vision_transformer.{PatchEmbed, Attention, Mlp, Block}  (call sites vit_mae.py:66,73-75;
               dreamvla_model.py:348-433; action_model/models.py:130,134)
  einops_exts  rearrange_many (perceiver_resampler.py:51)
  openai/CLIP  clip.load / clip.tokenize -> a random-weight text tower with CLIP ViT-B/32's text architecture
               (width 512, 12 layers, 8 heads, QuickGELU, causal mask, EOT pooling, text_projection)
"""
from __future__ import annotations

import os
import sys
import types
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("DVLA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


# ----------------------------------------------------------------------------------------------------------------------
# timm 0.9.16 restatement
# ----------------------------------------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0, proj_drop=0.0, norm_layer=nn.LayerNorm):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=self.attn_drop.p if self.training else 0.0)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None, bias=True,
                 drop=0.0, use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_norm=False, proj_drop=0.0, attn_drop=0.0,
                 init_values=None, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, mlp_layer=Mlp):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=proj_drop)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=proj_drop)
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


# ----------------------------------------------------------------------------------------------------------------------
# CLIP text tower stand-in (architecture of openai/CLIP ViT-B/32 text side, random weights)
# ----------------------------------------------------------------------------------------------------------------------
class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.attn_mask = attn_mask

    def forward(self, x):
        m = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=m)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class CLIPTextStandIn(nn.Module):
    """encode_text of openai/CLIP model.py (ViT-B/32 text hyper-parameters), random init with CLIP's init scheme."""

    def __init__(self, embed_dim=512, context_length=77, vocab_size=49408, width=512, heads=8, layers=12):
        super().__init__()
        self.context_length = context_length
        mask = torch.empty(context_length, context_length).fill_(float("-inf")).triu_(1)
        self.transformer = Transformer(width, layers, heads, attn_mask=mask)
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, width))
        self.ln_final = nn.LayerNorm(width)
        self.text_projection = nn.Parameter(torch.empty(width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=width ** -0.5)

    @property
    def dtype(self):
        return self.token_embedding.weight.dtype

    def encode_text(self, text):
        x = self.token_embedding(text).type(self.dtype)
        x = x + self.positional_embedding.type(self.dtype)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection
        return x


def _clip_load(name, device="cpu", jit=False, download_root=None):
    g = torch.random.get_rng_state()
    torch.manual_seed(20240607)
    model = CLIPTextStandIn().to(device)
    torch.random.set_rng_state(g)

    def preprocess(img):  # resize/normalise are data-side and out of scope
        return img

    return model, preprocess


def _clip_tokenize(texts, context_length=77, truncate=False):
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.int)
    for i, t in enumerate(texts):
        ids = [49406] + [1 + (hash_byte * 131 + j * 7) % 49404 for j, hash_byte in enumerate(t.encode()[:context_length - 2])] + [49407]
        out[i, :len(ids)] = torch.tensor(ids)
    return out


_installed = False


def install() -> None:
    """Register the shims and make the reference importable.  Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # transformers first: its lazy importer probes for a real `timm`, which must not see the stand-in half-built
    import transformers.pytorch_utils as pu
    for name in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):  # imported, never used (models/gpt2.py:14)
        if not hasattr(pu, name):
            setattr(pu, name, lambda *a, **k: None)
    import transformers.modeling_utils  # noqa: F401  (resolve lazies before `timm` appears in sys.modules)
    from transformers import GPT2Config  # noqa: F401
    from importlib.machinery import ModuleSpec
    # timm
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.PatchEmbed, vt.Block, vt.Attention, vt.Mlp = PatchEmbed, Block, Attention, Mlp
    vt.VisionTransformer = type("VisionTransformer", (nn.Module,), {})
    timm.models = timm_models
    timm_models.vision_transformer = vt
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", timm_models)
    sys.modules.setdefault("timm.models.vision_transformer", vt)
    # einops_exts
    from einops import rearrange
    ee = types.ModuleType("einops_exts")
    ee.rearrange_many = lambda ts, pattern, **kw: tuple(rearrange(t, pattern, **kw) for t in ts)
    sys.modules.setdefault("einops_exts", ee)
    # clip
    clip = types.ModuleType("clip")
    clip.load, clip.tokenize = _clip_load, _clip_tokenize
    sys.modules.setdefault("clip", clip)
    for mod in (timm, timm_models, vt, ee, clip):
        mod.__spec__ = ModuleSpec(mod.__name__, loader=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
