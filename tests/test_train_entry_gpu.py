"""train.py entry point on the GPU (single process): the reference's flags build the model, `train_one_epoch_calvin` runs on
synthetic collator tuples with gradient accumulation, `--cuda_graph` switches to graph replay after the first eager step, a
checkpoint round-trips through --resume_from_checkpoint INCLUDING the optimiser state (reference train.py:251-258)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

FLAGS = ["--finetune_type", "calvin", "--precision", "bf16", "--phase", "finetune", "--num_resampler_query", "16",
         "--num_obs_token_per_image", "9", "--transformer_layers", "2", "--hidden_dim", "1024", "--transformer_heads", "16",
         "--action_pred_steps", "3", "--sequence_length", "2", "--future_steps", "3", "--window_size", "5", "--obs_pred",
         "--loss_image", "--loss_action", "--use_dit_head", "--attn_implementation", "sdpa", "--batch_size", "1",
         "--learning_rate", "1e-5", "--weight_decay", "1e-4", "--lr_scheduler", "cosine", "--warmup_epochs", "0", "--seed", "7"]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def run(extra):
    import train
    from dreamvla_b200.utils.arguments_utils import get_parser
    args = get_parser().parse_args(FLAGS + extra)
    return train.main(args), args


def test_train_py_accumulation_graph_and_resume(dev, tmp_path):
    # 5 batches, accumulation 2: optimiser steps after batches 2, 4 and -- last-batch rule (train_utils.py:602-604) -- 5
    state, _ = run(["--synthetic_steps", "5", "--gradient_accumulation_steps", "2", "--cuda_graph"])
    assert state.total_micro == 5 and float(state.flat.step_count) == 3.0
    assert getattr(state.model, "_dvla_graphed_step", None) is not None, "--cuda_graph did not capture"
    assert all(torch.isfinite(v).all() for v in state.last_terms.values())
    assert float(state.flat.G.abs().max()) == 0.0                      # every accumulated gradient was consumed
    del state
    # checkpoint at the end of epoch 0, resume into epoch 1: weights, scheduler position and AdamW moments come back
    ck = str(tmp_path)
    s1, a1 = run(["--synthetic_steps", "2", "--num_epochs", "1", "--save_checkpoint", "--save_checkpoint_path", ck, "--run_name", "t",
                  "--start_save_checkpoint", "-1"])
    path = os.path.join(ck, "t", "0.pth")
    assert os.path.exists(path)
    saved = torch.load(path, map_location="cpu")
    assert set(saved) == {"epoch", "model_state_dict", "optimizer_state_dict", "lr_scheduler_state_dict"}
    assert all(k.startswith("module.") for k in saved["model_state_dict"])
    m1, v1, n1, P1 = s1.flat.m.clone(), s1.flat.v.clone(), float(s1.flat.step_count), s1.flat.P.clone()
    del s1
    # resume_from_epoch == num_epochs: builds the model, restores everything, trains nothing
    s2, _ = run(["--synthetic_steps", "1", "--num_epochs", "1", "--resume_from_checkpoint", path])
    assert s2.total_micro == 0 and float(s2.flat.step_count) == n1
    assert torch.equal(s2.flat.m, m1) and torch.equal(s2.flat.v, v1) and torch.equal(s2.flat.P, P1)
