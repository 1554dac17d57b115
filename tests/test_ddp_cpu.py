"""world_size-2 gloo test of the data-parallel exchange step (host-side logic of TrainStep.all_reduce_grads): one flat
SUM all-reduce, mean folded in afterwards as grad_scale = 1/world -- must equal the DDP gradient mean (train.py:173)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dreamvla_b200.utils.train_utils import all_reduce_flat
    g = torch.Generator().manual_seed(100 + rank)
    G = torch.randn(1000, generator=g).to(torch.bfloat16)       # this rank's local flat gradient
    local = G.float().clone()
    all_reduce_flat(G, world_size=world, group=None, comm_stream=None)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = torch.stack(gathered).mean(0)
    got = G.float() / world                                        # 1/world is applied by the AdamW kernel (grad_scale)
    ret[rank] = float((got - mean).abs().max())
    # every rank must hold the same reduced buffer -> identical optimiser updates
    ref = [torch.zeros(1000, dtype=torch.bfloat16) for _ in range(world)]
    dist.all_gather(ref, G)
    ret[10 + rank] = bool(all(torch.equal(ref[0], r) for r in ref))
    dist.destroy_process_group()


def test_flat_allreduce_is_ddp_mean():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] < 0.05, ret[r]          # bf16 sum rounding
        assert ret[10 + r]


def test_rank_seeds_shard_the_data():
    from dreamvla_b200.utils.train_utils import StepConfig, synthetic_batch
    cfg = StepConfig(sequence_length=2)
    a = synthetic_batch(cfg, 1, "cpu", seed=1234 + 0)
    b = synthetic_batch(cfg, 1, "cpu", seed=1234 + 1)
    assert not torch.equal(a["images_primary"], b["images_primary"])
    assert a["images_primary"].shape == (1, 5, 3, 224, 224) and a["text"].shape == (1, 77)


def test_gradient_segments_follow_backward_order():
    """The flat gradient buffer is all-reduced in pieces while backward runs (TrainStep._reduce_segment): segment 0 must only
    hold parameters used AFTER the backbone, segment k only backbone layers at or above the k-th cut, and nothing that feeds
    the backbone (projectors, resampler, query tokens, position embedding) may sit before the last segment -- it would be
    reduced before its gradient exists."""
    import re
    from dreamvla_b200.utils.train_utils import BACKBONE_CUTS, EARLY_GRAD_PREFIXES, backbone_cut_layers, grad_segment
    from tests import synth
    from tests.state_template import build_template
    assert backbone_cut_layers(24) == [18, 12, 6] and backbone_cut_layers(2) == [1] and backbone_cut_layers(1) == []
    for n_layers in (24, 2):
        cfg = dict(synth.CASES["calvin_allheads"], transformer_layers=n_layers)
        names = list(build_template(cfg).keys())
        cuts = backbone_cut_layers(n_layers)
        last = len(cuts) + 1
        seg = {n: grad_segment(n, n_layers) for n in names}
        feeds_backbone = ("perceiver_resampler.", "text_projector", "state_projector", "arm_state_encoder", "gripper_state_encoder",
                          "image_primary_projector", "image_wrist_projector", "cls_token_primary_projector",
                          "cls_token_wrist_projector", "obs_tokens", "depth_tokens", "dino_feat_tokens", "sam_feat_tokens",
                          "trajectory_tokens", "action_pred_token", "transformer_backbone_position_embedding",
                          "embedding_layer_norm", "vision_encoder.", "clip_model.")
        for n in names:
            if n.startswith(feeds_backbone):
                assert seg[n] == last, n
            m = re.match(r"transformer_backbone\.h\.(\d+)\.", n)
            if m:
                layer = int(m.group(1))
                want = next((k + 1 for k, c in enumerate(cuts) if layer >= c), last)
                assert seg[n] == want, n
            if seg[n] == 0:
                assert n.startswith(EARLY_GRAD_PREFIXES) and "tokens" not in n.split(".")[0].replace("mask_token", ""), n
        assert {0, 1, last} <= set(seg.values())
        assert seg["transformer_backbone.ln_f.weight"] == 1
    assert BACKBONE_CUTS >= 1


def test_on_grad_ready_fires_in_backward_completion_order():
    """ops.on_grad_ready marks: the callback of a tensor fires once everything downstream of it has run its backward --
    heads first, then the middle of the trunk -- which is the order the gradient segments are all-reduced in."""
    from dreamvla_b200.ops import on_grad_ready
    events = []
    w = [torch.randn(4, 4, requires_grad=True) for _ in range(5)]
    x = torch.randn(3, 4)
    h0 = x @ w[0]
    mid = torch.tanh(h0 @ w[1])
    on_grad_ready(mid, lambda: events.append(("mid", [wi.grad is not None for wi in w])))
    out = torch.tanh(mid @ w[2])
    on_grad_ready(out, lambda: events.append(("out", [wi.grad is not None for wi in w])))
    loss = (out @ w[3]).sum() + (out @ w[4]).pow(2).sum()       # two "heads" on the trunk output
    loss.backward()
    assert [e[0] for e in events] == ["out", "mid"]
    assert events[0][1] == [False, False, False, True, True]      # both heads done, trunk untouched
    assert events[1][1] == [False, False, True, True, True]       # second half of the trunk done
    assert all(wi.grad is not None for wi in w)


def test_prefetch_iterator_host_logic():
    """prefetch_to_device on a CPU device degenerates to a plain map: every batch, in order, exactly once (the CUDA path adds
    a copy stream and events around the same iteration order)."""
    from dreamvla_b200.utils.train_utils import prefetch_to_device
    seen = []

    def loader():
        for i in range(7):
            seen.append(i)
            yield {"x": torch.full((2,), float(i))}
    got = list(prefetch_to_device(loader(), "cpu", lambda hb: {k: v * 2 for k, v in hb.items()}))
    assert [int(b["x"][0]) for b in got] == [0, 2, 4, 6, 8, 10, 12] and seen == list(range(7))
    assert all(b["x"].dtype == torch.bfloat16 for b in got)           # floating-point entries arrive in the step's dtype
    assert list(prefetch_to_device(iter(()), "cpu")) == []
