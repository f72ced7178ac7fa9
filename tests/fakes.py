"""Deterministic stand-in producers (encoders / tokenisers / processors) for tests.

The encoders are OUT of the hot path (they stay PyTorch producers), and no
weights exist offline, so both the golden generator (which drives the REAL reference
methods) and the drop-in tests (which drive mmf_b200) plug in the same fakes: an image
or text carries an integer id, and the fake encoder returns row `id` of a seeded table.
That makes the embeddings / head scores reaching the hot path bit-identical on both sides.
"""
from __future__ import annotations

import math
import re

import numpy as np
import torch
import torch.nn as nn
from PIL import Image


def image_for_id(i: int) -> Image.Image:
    """2x2 RGB image whose pixel encodes a 24-bit id."""
    px = (int(i) >> 16 & 255, int(i) >> 8 & 255, int(i) & 255)
    return Image.new("RGB", (2, 2), px)


def id_of_image(img: Image.Image) -> int:
    r, g, b = img.convert("RGB").getpixel((0, 0))
    return (r << 16) | (g << 8) | b


def text_for_id(i: int) -> str:
    return f"caption #{int(i)}"


def id_of_text(s: str) -> int:
    m = re.search(r"#(\d+)", s)
    if m is None:
        raise KeyError(f"fake encoder: no id in {s!r}")
    return int(m.group(1))


class _Batch(dict):
    """What a HF processor/tokeniser returns: a dict with .to(device)."""

    def to(self, device):
        return _Batch({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})


class FakeClipProcessor:
    def __call__(self, text=None, images=None, return_tensors="pt", padding=True, truncation=False, max_length=None):
        out = _Batch()
        if text is not None:
            out["input_ids"] = torch.tensor([[id_of_text(t)] for t in text], dtype=torch.long)
        if images is not None:
            imgs = images if isinstance(images, (list, tuple)) else [images]
            out["pixel_values"] = torch.tensor([[id_of_image(im)] for im in imgs], dtype=torch.long)
        return out


class _ClipOut:
    def __init__(self, text_embeds, image_embeds):
        self.text_embeds = text_embeds
        self.image_embeds = image_embeds


class FakeClipModel(nn.Module):
    """Returns rows of seeded tables.  get_*_features return plain tensors (the
    transformers-4 behaviour the reference was written against)."""

    def __init__(self, image_table: np.ndarray, text_table: np.ndarray):
        super().__init__()
        self.register_buffer("image_table", torch.as_tensor(image_table, dtype=torch.float32))
        self.register_buffer("text_table", torch.as_tensor(text_table, dtype=torch.float32))

    def get_image_features(self, pixel_values=None, **kw):
        return self.image_table[pixel_values[:, 0].to(self.image_table.device)]

    def get_text_features(self, input_ids=None, **kw):
        return self.text_table[input_ids[:, 0].to(self.text_table.device)]

    def forward(self, input_ids=None, pixel_values=None, **kw):
        return _ClipOut(self.get_text_features(input_ids), self.get_image_features(pixel_values))


class FakeTokenizer:
    def __call__(self, text, return_tensors="pt", max_length=512, truncation=True, padding=True):
        texts = [text] if isinstance(text, str) else list(text)
        return _Batch(input_ids=torch.tensor([[id_of_text(t)] for t in texts], dtype=torch.long),
                      attention_mask=torch.ones(len(texts), 1, dtype=torch.long))


def _logit(p: float) -> float:
    p = min(max(float(p), 1e-6), 1 - 1e-6)
    return math.log(p / (1 - p))


class FakeDetector(nn.Module):
    """Planted head scores + a REAL fusion_layer with the reference's architecture
    (misinfo_forensics.py:83-90)."""

    def __init__(self, ai_scores, misinfo_scores, deepfake_scores, fusion_seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(fusion_seed)
        self.fusion_layer = nn.Sequential(
            nn.Linear(5, 64), nn.ReLU(), nn.Dropout(0.2),
            nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 2))
        with torch.no_grad():
            for p in self.fusion_layer.parameters():
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.8)
        self.register_buffer("ai", torch.tensor([_logit(p) for p in ai_scores]))
        self.register_buffer("mis", torch.tensor([_logit(p) for p in misinfo_scores]))
        self.register_buffer("deep", torch.tensor([_logit(p) for p in deepfake_scores]))

    @staticmethod
    def _two(logit_col):
        return torch.stack([torch.zeros_like(logit_col), logit_col], dim=1)

    def forward_text(self, input_ids, attention_mask):
        i = input_ids[:, 0]
        return self._two(self.ai[i]), self._two(self.mis[i])

    def forward_image(self, image_tensor):
        # image_tensor is the EfficientNet-normalised (B,3,224,224) image; recover the id
        mean = torch.tensor([0.485, 0.456, 0.406], device=image_tensor.device)
        std = torch.tensor([0.229, 0.224, 0.225], device=image_tensor.device)
        px = (image_tensor[:, :, 0, 0] * std + mean) * 255.0
        px = px.round().long()
        i = (px[:, 0] << 16) | (px[:, 1] << 8) | px[:, 2]
        return self._two(self.deep[i])

    def forward_fusion(self, scores_tensor):
        return self.fusion_layer(scores_tensor)


class FakeCv2:
    """Minimal cv2 surface used by analyze_video (misinfo_forensics.py:500-545)."""
    CAP_PROP_FPS = 5
    COLOR_BGR2RGB = 4

    class VideoCapture:
        def __init__(self, path):
            # "fake://fps=2;ids=3,4,5,6"
            m = re.match(r"fake://fps=([\d.]+);ids=([\d,]*)", str(path))
            self._ok = m is not None
            self._fps = float(m.group(1)) if m else 0.0
            self._ids = [int(x) for x in m.group(2).split(",") if x] if m else []
            self._pos = 0

        def isOpened(self):
            return self._ok

        def get(self, prop):
            return self._fps

        def read(self):
            if self._pos >= len(self._ids):
                return False, None
            i = self._ids[self._pos]
            self._pos += 1
            frame = np.zeros((2, 2, 3), np.uint8)
            frame[...] = (i & 255, i >> 8 & 255, i >> 16 & 255)      # BGR
            return True, frame

        def release(self):
            pass

    @staticmethod
    def cvtColor(frame, code):
        return frame[..., ::-1].copy()
