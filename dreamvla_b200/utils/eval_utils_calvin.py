"""Action-inference wrapper -- mirror of reference utils/eval_utils_calvin.py:48-147 (`ModelWrapper.step`) and of the
LIBERO variant's window handling (utils/eval_utils_libero.py:94-179), without the simulators (out of scope).

Semantics kept: growing-then-sliding window of `history_len` frames, window padded by repeating the last frame, the
instruction tokens repeated over the window, full-window `mode='test'` forward every env step, gripper thresholding and
selection of row `num_step-1` (or the last row once the window is full).  Inputs are already-preprocessed tensors
(CLIP image preprocessing / tokenisation are data-side).
"""
from __future__ import annotations

from collections import deque

import torch


class ModelWrapper:
    def __init__(self, model, cast_dtype=torch.bfloat16, history_len=10, action_pred_steps=3, device="cuda",
                 use_cuda_graph=True):
        self.model = model.module if hasattr(model, "module") else model
        self.cast_type = cast_dtype
        self.history_len = history_len
        self.action_pred_steps = action_pred_steps
        self.device = device
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.reset()

    def _forward(self, image_primary, image_wrist, state, text_token, sample_noise):
        """Full-window mode='test' forward.  With use_cuda_graph the ~2.5 k kernel launches of one action (ViT on 2*S frames,
        resampler, 24-layer backbone, 10 DDIM steps x DiT-B) are captured once and replayed as ONE graph launch."""
        if not self.use_cuda_graph:
            return self.model(image_primary=image_primary, image_wrist=image_wrist, state=state, text_token=text_token,
                              action=None, mode="test", sample_noise=sample_noise)
        if self._graph is None:
            n = self.history_len
            self._static = dict(image_primary=torch.zeros_like(image_primary), image_wrist=torch.zeros_like(image_wrist),
                                state=torch.zeros_like(state), text_token=torch.zeros(1, 77, dtype=text_token.dtype, device=self.device),
                                noise=torch.zeros(n, self.action_pred_steps, 7, device=self.device))
            st = self._static

            def run():
                return self.model(image_primary=st["image_primary"], image_wrist=st["image_wrist"], state=st["state"],
                                  text_token=st["text_token"].unsqueeze(1).expand(1, n, 77), action=None, mode="test",
                                  sample_noise=st["noise"])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._out = run()
        st = self._static
        st["image_primary"].copy_(image_primary)
        st["image_wrist"].copy_(image_wrist)
        st["state"].copy_(state)
        st["text_token"].copy_(text_token[:, 0])
        if sample_noise is None:
            st["noise"].normal_()
        else:
            st["noise"].copy_(sample_noise)
        self._graph.replay()
        return self._out

    def reset(self):
        self.img_queue = deque(maxlen=self.history_len)
        self.gripper_queue = deque(maxlen=self.history_len)
        self.state_queue = deque(maxlen=self.history_len)
        self.text_token = None

    @torch.no_grad()
    def step(self, image_static, image_gripper, robot_obs, text_tokens, sample_noise=None):
        """image_* [3,224,224] preprocessed, robot_obs [15], text_tokens int [77] -> action [7] (fp16 numpy like the reference)."""
        dev = self.device
        self.img_queue.append(image_static.to(dev, self.cast_type).view(1, 1, 3, 224, 224))
        self.gripper_queue.append(image_gripper.to(dev, self.cast_type).view(1, 1, 3, 224, 224))
        st = robot_obs.to(dev, self.cast_type).view(1, 1, -1)
        self.state_queue.append(torch.cat([st[..., :6], st[..., -1:]], dim=-1))
        if self.text_token is None:                              # frozen for the episode (eval_utils_calvin.py:110-113)
            self.text_token = text_tokens.to(dev).view(1, 1, 77).expand(1, self.history_len, 77)
        image_primary = torch.cat(list(self.img_queue), dim=1)
        image_wrist = torch.cat(list(self.gripper_queue), dim=1)
        state = torch.cat(list(self.state_queue), dim=1)
        num_step = image_primary.shape[1]
        if num_step < self.history_len:
            pad = self.history_len - num_step
            image_primary = torch.cat([image_primary, image_primary[:, -1:].expand(-1, pad, -1, -1, -1)], dim=1)
            image_wrist = torch.cat([image_wrist, image_wrist[:, -1:].expand(-1, pad, -1, -1, -1)], dim=1)
            state = torch.cat([state, state[:, -1:].expand(-1, pad, -1)], dim=1)
        out = self._forward(image_primary, image_wrist, state, self.text_token, sample_noise)
        arm_action, gripper_action = out[0], out[1]
        action = torch.cat((arm_action[0, :, 0, :].float(), (gripper_action[0, :, 0, :] > 0.5).float()), dim=-1)
        action[:, -1] = (action[:, -1] - 0.5) * 2
        row = num_step - 1 if num_step < self.history_len else -1
        return action[row].to(torch.float16).cpu().numpy()
