"""state_dict template (key names, shapes, fixed buffers) built from the PRODUCT model's constructor on CPU.

Used off-box in place of the reference's state_dict; tests/test_oracle_cpu.py::test_state_template_matches_reference_keys
checks in the build container that it equals the reference's state_dict layout."""
import functools

import torch

from tests import synth


@functools.lru_cache(maxsize=4)
def _build(cfg_items):
    from dreamvla_b200.models import DreamVLA
    cfg = dict(cfg_items)
    torch.manual_seed(0)
    m = DreamVLA(finetune_type="calvin", clip_device="cpu", vit_checkpoint_path=None, **synth.ctor_kwargs(cfg))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def build_template(cfg):
    return _build(tuple(sorted((k, v) for k, v in cfg.items())))
