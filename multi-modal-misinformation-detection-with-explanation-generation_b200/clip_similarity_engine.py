"""Drop-in for the reference's clip_similarity_engine.CLIPSimilarityEngine: same
constructor, methods, return values and error behaviour; the CLIP encoder stays a PyTorch
producer and the normalise + cosine + Match rule (clip_similarity_engine.py:103-111) runs
in libmmf_b200's cosine kernel."""
from __future__ import annotations

import os

import torch
from PIL import Image

from .engine import Engine, MATCH_THRESHOLD


class CLIPSimilarityEngine:
    def __init__(self, model_name="openai/clip-vit-base-patch32", threshold=MATCH_THRESHOLD, *, model=None,
                 processor=None, engine: Engine = None, device=None):
        """model/processor may be injected (offline use); otherwise they are loaded with
        from_pretrained exactly like the reference."""
        print(f"Loading CLIP model: {model_name}...")
        try:
            if model is None or processor is None:
                from transformers import CLIPModel, CLIPProcessor
                model = model or CLIPModel.from_pretrained(model_name)
                processor = processor or CLIPProcessor.from_pretrained(model_name)
            self.engine = engine or Engine(device or "cuda")
            self.model, self.processor, self.threshold = model, processor, threshold
            self.device = str(self.engine.device)
            self.model.to(self.device)
            print(f"Model loaded successfully on {self.device}")
        except Exception as e:
            raise RuntimeError(f"Failed to load CLIP model: {str(e)}")

    def load_image(self, image_path):
        if not os.path.exists(image_path):
            raise FileNotFoundError(f"Image file not found: {image_path}")
        try:
            image = Image.open(image_path)
            return image if image.mode == "RGB" else image.convert("RGB")
        except Exception as e:
            raise ValueError(f"Failed to load image from {image_path}: {str(e)}")

    def calculate_similarity(self, image_path, text):
        """-> (cosine similarity, 'Match' | 'Mismatch')."""
        try:
            image = self.load_image(image_path)
            if not text or not isinstance(text, str):
                raise ValueError("Text input must be a non-empty string")
            inputs = self.processor(text=[text], images=image, return_tensors="pt", padding=True)
            inputs = {k: v.to(self.device) for k, v in inputs.items()}
            with torch.no_grad():
                outputs = self.model(**inputs)
            sim, match = self.engine.cosine_pairs(outputs.image_embeds, outputs.text_embeds, self.threshold)
            similarity = sim.item()
            return similarity, ("Match" if bool(match.item()) else "Mismatch")
        except (FileNotFoundError, ValueError):
            raise
        except Exception as e:
            raise RuntimeError(f"Error calculating similarity: {str(e)}")

    def analyze_with_explanation(self, image_path, text):
        try:
            similarity, label = self.calculate_similarity(image_path, text)
            return {"image_path": image_path, "text": text, "similarity_score": round(similarity, 4),
                    "label": label, "explanation": self._generate_explanation(similarity, label)}
        except Exception as e:
            return {"image_path": image_path, "text": text, "error": str(e)}

    _TIERS = {
        "Match": ((0.7, "Strong match detected (score: {s:.4f}). The image and text are highly consistent."),
                  (0.5, "Moderate match detected (score: {s:.4f}). The image and text show reasonable alignment."),
                  (None, "Weak match detected (score: {s:.4f}). The image and text are barely above the threshold.")),
    }

    def _generate_explanation(self, similarity, label):
        if label == "Match":
            for floor, msg in self._TIERS["Match"]:
                if floor is None or similarity >= floor:
                    return msg.format(s=similarity)
        if similarity < 0.1:
            return (f"Strong mismatch detected (score: {similarity:.4f}). "
                    "The image and text appear completely unrelated.")
        return (f"Mismatch detected (score: {similarity:.4f}). "
                "The image and text show inconsistencies that may indicate misinformation.")
