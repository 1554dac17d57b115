from .dreamvla_model import DreamVLA, generate_attention_mask  # noqa: F401

DreamVLAModel = DreamVLA  # BASELINE.json's name for the class
