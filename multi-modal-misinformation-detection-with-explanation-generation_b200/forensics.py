"""Drop-in for the reference's misinfo_forensics.MisinfoForensics.

Same constructor arguments, methods, return dictionaries and error behaviour
(misinfo_forensics.py:111-927).  The RoBERTa / EfficientNet / CLIP encoders stay PyTorch
producers; everything downstream of their outputs -- caption/image cosine (:399-404),
Truth-Vault search + discrepancy rule (:438-464), caption/headline similarity (:481-484),
fusion judge + verdict (:587-608) and the batched forms of the same -- runs in
libmmf_b200.so on the B200.  There is no CPU path: construction fails without an sm_100
device.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn
from PIL import Image

from .engine import Engine, VAULT_THRESHOLD
from .pipeline import score_batch
from .vault import TruthVault, load_vault_file, read_vault_dict

SCORE_ORDER = ("ai_score", "misinfo_score", "deepfake_score", "clip_similarity", "vault_discrepancy")


class MultiModalMisinfoDetector(nn.Module):
    """The producer network (misinfo_forensics.py:43-108); attribute names are the .pth
    state-dict contract ('roberta', 'ai_head', 'misinfo_head', 'efficientnet', 'fusion_layer')."""

    def __init__(self, roberta_model_name: str = "roberta-base", roberta=None):
        super().__init__()
        from torchvision import models
        if roberta is None:
            from transformers import RobertaModel
            roberta = RobertaModel.from_pretrained(roberta_model_name)
        self.roberta = roberta
        hidden = self.roberta.config.hidden_size

        def head():
            return nn.Sequential(nn.Linear(hidden, 256), nn.ReLU(), nn.Dropout(0.3), nn.Linear(256, 2))
        self.ai_head, self.misinfo_head = head(), head()
        self.efficientnet = models.efficientnet_b0(weights=None)
        self.efficientnet.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(1280, 2))
        self.fusion_layer = nn.Sequential(nn.Linear(5, 64), nn.ReLU(), nn.Dropout(0.2),
                                          nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 2))

    def forward_text(self, input_ids, attention_mask):
        cls = self.roberta(input_ids=input_ids, attention_mask=attention_mask).last_hidden_state[:, 0, :]
        return self.ai_head(cls), self.misinfo_head(cls)

    def forward_image(self, image_tensor):
        return self.efficientnet(image_tensor)

    def forward_fusion(self, scores_tensor):
        return self.fusion_layer(scores_tensor)


def _features(x):
    """transformers >= 5 returns BaseModelOutputWithPooling from get_*_features; the projected
    embedding is .pooler_output."""
    return x if torch.is_tensor(x) else x.pooler_output


def _load_checked(module: nn.Module, state: dict, what: str) -> None:
    """load_state_dict(strict=False) that SAYS when nothing matched (strict=False never raises for that)."""
    res = module.load_state_dict(state, strict=False)
    own = set(module.state_dict().keys())
    if not state or own.issubset(set(res.missing_keys)):
        print(f"  ⚠ {what}: no parameter of the checkpoint matched -- weights stay as initialised")


def load_individual_weights(detector: nn.Module, ai_head_weights: str, misinfo_head_weights: str,
                            efficientnet_weights: str) -> None:
    """Per-branch fallback of the reference (misinfo_forensics.py:260-304): the head checkpoints are
    {'model_state_dict': {'ai_head.0.weight', ...}, 'epoch', ...} -- filter by branch name, strip the prefix;
    EfficientNet is either that form (prefix 'efficientnet.') or a raw state_dict (whole net, else classifier only).
    (The reference also tries `clip_detective_best.pth` here, but at that point its CLIP model does not exist yet,
    the attempt always ends in its except branch and from_pretrained follows -- so nothing to mirror.)"""
    for path, branch, target in ((ai_head_weights, "ai_head", detector.ai_head),
                                 (misinfo_head_weights, "misinfo_head", detector.misinfo_head)):
        if os.path.exists(path):
            print(f"Loading {branch} from {path}...")
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            state = {k.replace(branch + ".", ""): v for k, v in ckpt["model_state_dict"].items() if branch in k}
            _load_checked(target, state, path)
            print(f"  ✓ Loaded from epoch {ckpt.get('epoch', 'N/A')}")
    if os.path.exists(efficientnet_weights):
        print(f"Loading EfficientNet from {efficientnet_weights}...")
        ckpt = torch.load(efficientnet_weights, map_location="cpu", weights_only=False)
        if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
            state = {k.replace("efficientnet.", ""): v for k, v in ckpt["model_state_dict"].items() if "efficientnet" in k}
            _load_checked(detector.efficientnet, state, efficientnet_weights)
            print(f"  ✓ Loaded from epoch {ckpt.get('epoch', 'N/A')}")
        else:
            try:
                _load_checked(detector.efficientnet, ckpt, efficientnet_weights)
                print("  ✓ Loaded weights successfully")
            except RuntimeError:
                cls = {k: v for k, v in ckpt.items() if "classifier" in k}
                if cls:
                    detector.efficientnet.classifier.load_state_dict(cls, strict=False)
                    print("  ✓ Loaded classifier head")


class MisinfoForensics:
    def __init__(
        self,
        fusion_weights: str = "forensics_master_final.pth",
        ai_head_weights: str = "ai_head_best.pth",
        misinfo_head_weights: str = "roberta_detective_best.pth",
        efficientnet_weights: str = "efficientnet_cifake_best.pth",
        clip_model_dir: str = r"C:\Users\Lenovo\OneDrive\Desktop\hack\models\clip-vit-b32",
        clip_weights: str = "clip_detective_best.pth",
        faiss_index_path: str = "guardian_embeddings.pkl",
        gemini_api_key: Optional[str] = None,
        device: str = "cuda",
        *,
        detector: Optional[nn.Module] = None,
        roberta_tokenizer=None,
        clip_model: Optional[nn.Module] = None,
        clip_processor=None,
        vault: Optional[dict] = None,
        vault_mode: str = "fp32",
        explainer=None,
        engine: Optional[Engine] = None,
        dedup_clip_encode: bool = True,
    ):
        """Positional arguments are the reference's.  The keyword-only ones let a caller inject
        already-built producers / an in-memory vault dict (offline use, tests) and choose the
        resident vault precision; `explainer(all_scores, vault_matches) -> str` replaces the
        Gemini call (out of scope here), default is the reference's rule-based text."""
        self.engine = engine or Engine(device)
        self.device = self.engine.device
        print(f"Using device: {self.device}")
        self.gemini_available = False
        self.explainer = explainer
        # analyze() encodes the image ONCE for both the consistency and the vault step (the reference
        # runs the CLIP image tower twice, misinfo_forensics.py:395 and :438); results are identical
        self.dedup_clip_encode = dedup_clip_encode

        if roberta_tokenizer is None:
            from transformers import RobertaTokenizer
            roberta_tokenizer = RobertaTokenizer.from_pretrained("roberta-base")
        self.roberta_tokenizer = roberta_tokenizer

        if detector is None:
            detector = MultiModalMisinfoDetector("roberta-base")
            self._load_detector_weights(detector, fusion_weights, ai_head_weights, misinfo_head_weights,
                                        efficientnet_weights)
        self.detector = detector.to(self.device).eval()
        self.reload_fusion()

        if clip_model is None or clip_processor is None:
            from transformers import CLIPModel, CLIPProcessor
            clip_processor = clip_processor or CLIPProcessor.from_pretrained(clip_model_dir)
            clip_model = clip_model or CLIPModel.from_pretrained(clip_model_dir)
        self.clip_processor = clip_processor
        self.clip_model = clip_model.to(self.device).eval()

        # Truth Vault: host attributes stay as in the reference, the rows are uploaded once
        self.vault_loaded = False
        self.vault_data, self.vault_embeddings, self.vault_metadata, self.vault = None, None, None, None
        if vault is not None:
            self.vault_data = vault
            self.vault_embeddings, self.vault_metadata = read_vault_dict(vault)
        elif os.path.exists(faiss_index_path):
            print(f"\nLoading Truth Vault from {faiss_index_path}...")
            self.vault_data, self.vault_embeddings, self.vault_metadata = load_vault_file(faiss_index_path)
        else:
            print(f"⚠ Truth Vault not found: {faiss_index_path}")
        if self.vault_embeddings is not None:
            self.vault = TruthVault(self.engine, self.vault_embeddings, self.vault_metadata, mode=vault_mode)
            self.vault_loaded = True
            print(f"  ✓ Loaded {len(self.vault_metadata)} verified articles")
        elif self.vault_data is not None:
            print("  ⚠ Unknown database format")

        from torchvision import transforms
        self.efficientnet_transform = transforms.Compose([
            transforms.Resize((224, 224)), transforms.ToTensor(),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])

    # ------------------------------------------------------------------ weights
    def _load_detector_weights(self, detector, fusion_weights, ai_w, mis_w, eff_w):
        """full_model_state_dict first (misinfo_forensics.py:175-197), else per-branch files (:260-304)."""
        if os.path.exists(fusion_weights):
            try:
                ckpt = torch.load(fusion_weights, map_location="cpu", weights_only=False)
                detector.load_state_dict(ckpt["full_model_state_dict"], strict=False)
                print(f"  ✓ Loaded complete integrated model from {fusion_weights}")
                return
            except Exception as e:
                print(f"  ⚠ Error loading fusion weights: {e}")
                print("  Falling back to individual model loading...")
        load_individual_weights(detector, ai_w, mis_w, eff_w)

    def reload_fusion(self):
        """Re-read detector.fusion_layer into the library (the fusion trainer mutates those
        weights in place, train_fusion_judge.py:144-233)."""
        self.engine.fusion_load(self.detector.fusion_layer.state_dict())

    def _to_pil_image(self, image_or_path: Union[str, Image.Image]) -> Image.Image:
        if isinstance(image_or_path, Image.Image):
            return image_or_path.convert("RGB")
        return Image.open(str(image_or_path)).convert("RGB")

    # ------------------------------------------------------------------ producers (PyTorch)
    def analyze_text(self, text: str) -> Dict[str, float]:
        inputs = self.roberta_tokenizer(text, return_tensors="pt", max_length=512, truncation=True,
                                        padding=True).to(self.device)
        with torch.no_grad():
            ai_logits, mis_logits = self.detector.forward_text(inputs["input_ids"], inputs["attention_mask"])
            both = torch.stack([torch.softmax(ai_logits, dim=1)[0, 1], torch.softmax(mis_logits, dim=1)[0, 1]]).tolist()
        return {"ai_score": both[0], "misinfo_score": both[1]}

    def analyze_image(self, image_path: Union[str, Image.Image]) -> Dict[str, float]:
        image = self._to_pil_image(image_path)
        t = self.efficientnet_transform(image).unsqueeze(0).to(self.device)
        with torch.no_grad():
            p = torch.softmax(self.detector.forward_image(t), dim=1)[0, 1].item()
        return {"deepfake_score": p}

    def analyze_text_batch(self, texts: Sequence[str], batch_size: int = 64) -> np.ndarray:
        """analyze_text for many texts: (n,2) [ai_score, misinfo_score], the RoBERTa forward run on padded batches
        (the producers are out of this repo's scope, but a per-sample loop around them would make the batched hot
        path pointless).  batch_size=1 reproduces analyze_text's own forward exactly."""
        out = np.zeros((len(texts), 2), np.float32)
        for b0 in range(0, len(texts), max(1, batch_size)):
            chunk = list(texts[b0:b0 + max(1, batch_size)])
            inputs = self.roberta_tokenizer(chunk if len(chunk) > 1 else chunk[0], return_tensors="pt", max_length=512,
                                            truncation=True, padding=True).to(self.device)
            with torch.no_grad():
                ai_logits, mis_logits = self.detector.forward_text(inputs["input_ids"], inputs["attention_mask"])
                both = torch.stack([torch.softmax(ai_logits, dim=1)[:, 1], torch.softmax(mis_logits, dim=1)[:, 1]], dim=1)
            out[b0:b0 + len(chunk)] = both.float().cpu().numpy()
        return out

    def analyze_image_batch(self, images: Sequence[Image.Image], batch_size: int = 64) -> np.ndarray:
        """analyze_image for many (already opened) images: (n,) deepfake_score, EfficientNet on stacked batches."""
        out = np.zeros(len(images), np.float32)
        for b0 in range(0, len(images), max(1, batch_size)):
            chunk = images[b0:b0 + max(1, batch_size)]
            t = torch.stack([self.efficientnet_transform(im) for im in chunk]).to(self.device)
            with torch.no_grad():
                out[b0:b0 + len(chunk)] = torch.softmax(self.detector.forward_image(t), dim=1)[:, 1].float().cpu().numpy()
        return out

    def _clip_image_embed(self, images: Sequence[Image.Image]) -> torch.Tensor:
        inputs = self.clip_processor(images=list(images) if len(images) > 1 else images[0], return_tensors="pt").to(self.device)
        with torch.no_grad():
            return _features(self.clip_model.get_image_features(**inputs))

    def _clip_text_embed(self, texts: Sequence[str], truncation: bool = True) -> torch.Tensor:
        """truncation=True is the caption/headline step (misinfo_forensics.py:471-476); the caption/image consistency
        step does NOT truncate (:386-391: a caption over 77 tokens raises there, and so it does here)."""
        inputs = self.clip_processor(text=list(texts), return_tensors="pt", padding=True, truncation=truncation).to(self.device)
        with torch.no_grad():
            return _features(self.clip_model.get_text_features(**inputs))

    # ------------------------------------------------------------------ hot path
    def analyze_consistency(self, text: str, image_path: Union[str, Image.Image]) -> Dict[str, float]:
        image = self._to_pil_image(image_path)
        inputs = self.clip_processor(text=[text], images=image, return_tensors="pt", padding=True).to(self.device)
        with torch.no_grad():
            out = self.clip_model(**inputs)
        return {"clip_similarity": self.engine.cosine_pairs(out.text_embeds, out.image_embeds).item()}

    def search_vault(self, image_path: Union[str, Image.Image], user_caption: str = None, top_k: int = 5) -> Dict:
        if not self.vault_loaded:
            return {"vault_discrepancy": 0.0, "matches": [], "vault_available": False, "text_similarity": 0.0}
        emb = self._clip_image_embed([self._to_pil_image(image_path)])
        return self._vault_results(emb, [user_caption], top_k)[0]

    def _vault_results(self, image_embeds: torch.Tensor, captions: Sequence[Optional[str]], top_k: int) -> List[Dict]:
        """Batched search_vault body: one device search for all rows, one D2H, then the
        host-side match records; the caption/headline cosine only for rows over the threshold."""
        scores, rows, disc = self.vault.search(image_embeds, top_k, VAULT_THRESHOLD)
        scores, rows, disc = scores.cpu().numpy(), rows.cpu().numpy(), disc.cpu().numpy()
        results, need = [], []
        for i in range(len(captions)):
            matches = self.vault.matches(scores[i], rows[i])
            d = float(disc[i])
            results.append({"vault_discrepancy": d, "matches": matches, "vault_available": True, "text_similarity": 0.0})
            if captions[i] and d != 0.0 and matches:
                need.append(i)
        if need:
            emb = self._clip_text_embed([t for i in need for t in (captions[i], results[i]["matches"][0]["title"])])
            sims = self.engine.cosine_pairs(emb[0::2], emb[1::2]).tolist()
            for i, s in zip(need, sims):
                results[i]["text_similarity"] = float(s)
        return results

    def fusion_verdict(self, scores: Dict[str, float]) -> Dict:
        x = torch.tensor([[scores.get(k, 0.0) for k in SCORE_ORDER]], dtype=torch.float32)
        probs, verdict, conf = self.engine.fusion_forward(x)
        real, fake = probs[0].tolist()
        label = int(verdict.item())
        return {"verdict": label, "confidence": fake if label == 1 else real,
                "fake_probability": fake, "real_probability": real}

    # ------------------------------------------------------------------ video (batched consumer)
    def analyze_video(self, video_path: str, text: Optional[str] = None, max_frames: int = 12,
                      stride_seconds: float = 1.0) -> Dict:
        """Frame sampling as misinfo_forensics.py:493-545; the sampled frames then go through
        the hot path as ONE batch instead of one vault pass per frame."""
        try:
            import cv2
        except Exception as e:
            raise RuntimeError("opencv-python is required for video analysis. Install with: pip install opencv-python") from e
        cap = cv2.VideoCapture(video_path)
        if not cap.isOpened():
            raise RuntimeError(f"Could not open video: {video_path}")
        fps = cap.get(cv2.CAP_PROP_FPS)
        if not fps or fps <= 0:
            fps = 25.0
        stride = max(1, int(round(fps * max(0.1, float(stride_seconds)))))
        frames: List[Image.Image] = []
        idx = 0
        while len(frames) < max_frames:
            ok, frame = cap.read()
            if not ok:
                break
            if idx % stride == 0:
                frames.append(Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)))
            idx += 1
        cap.release()
        if not frames:
            raise RuntimeError("No frames could be read from the video.")

        deepfake = [float(self.analyze_image(f)["deepfake_score"]) for f in frames]
        clip_sims: List[float] = []
        if text:
            inputs = self.clip_processor(text=[text], images=frames if len(frames) > 1 else frames[0],
                                         return_tensors="pt", padding=True).to(self.device)
            with torch.no_grad():
                out = self.clip_model(**inputs)
            t = out.text_embeds.expand(out.image_embeds.shape[0], -1)
            clip_sims = [float(s) for s in self.engine.cosine_pairs(t, out.image_embeds).tolist()]
        best = {"vault_discrepancy": 0.0, "matches": [], "vault_available": self.vault_loaded, "text_similarity": 0.0}
        best_frame = None
        if self.vault_loaded:
            per_frame = self._vault_results(self._clip_image_embed(frames), [text] * len(frames), 5)
            for f, v in zip(frames, per_frame):             # first strictly-larger discrepancy wins (:554)
                if float(v["vault_discrepancy"]) > float(best["vault_discrepancy"]):
                    best, best_frame = v, f
        return {"deepfake_score": float(np.mean(deepfake)),
                "clip_similarity": float(np.mean(clip_sims)) if clip_sims else 0.0,
                "vault_discrepancy": float(best.get("vault_discrepancy", 0.0)),
                "text_similarity": float(best.get("text_similarity", 0.0)),
                "vault_matches": best.get("matches", []), "best_frame": best_frame}

    # ------------------------------------------------------------------ explanation (not hot path)
    def generate_gemini_explanation(self, all_scores: Dict, vault_matches: list) -> str:
        if self.explainer is not None:
            try:
                text = self.explainer(all_scores, vault_matches)
                if text:
                    return str(text).strip()
            except Exception as e:
                print(f"  ⚠ explainer failed: {e}\n  Falling back to rule-based explanation")
        return self._generate_fallback_explanation(all_scores, vault_matches)

    def _generate_fallback_explanation(self, s: Dict, vault_matches: list) -> str:
        """Rule-based summary, same sentences as misinfo_forensics.py:742-765."""
        head = f"This content is classified as {'FAKE' if s['verdict'] == 1 else 'REAL'}"
        rules = (
            (s["vault_discrepancy"] > 0.7, lambda: "Our database found this image was previously published in a different "
             f"context (\"{vault_matches[0]['title']}\"), suggesting potential misuse."),
            (s["deepfake_score"] > 0.7, lambda: "The image shows strong signs of digital manipulation "
             f"(deepfake probability: {s['deepfake_score']:.1%})."),
            (s["ai_score"] > 0.7, lambda: "The text exhibits characteristics typical of AI-generated content."),
            (s["misinfo_score"] > 0.7, lambda: "The text uses language patterns commonly associated with misinformation."),
            (s["clip_similarity"] < 0.3, lambda: "The image and caption show poor alignment, suggesting potential mismatching."),
        )
        for hit, tail in rules:
            if hit:
                return f"{head}. {tail()}"
        return (f"{head} with {s['confidence']:.1%} confidence. Multiple signals from text analysis, "
                "image forensics, and database checks support this assessment.")

    # ------------------------------------------------------------------ orchestration
    @staticmethod
    def _fallback_verdict(scores: Dict, has_text: bool, has_visual: bool) -> Dict:
        """Missing-modality rule, misinfo_forensics.py:884-899 (host scalars, no kernel needed)."""
        if has_text and not has_visual:
            fake = float(scores.get("misinfo_score", 0.0))
        elif has_visual and not has_text:
            fake = float(max(scores.get("deepfake_score", 0.0), scores.get("vault_discrepancy", 0.0)))
        else:
            fake = 0.5
        fake = max(0.0, min(1.0, fake))
        real = 1.0 - fake
        label = 1 if fake > 0.5 else 0
        return {"verdict": label, "confidence": fake if label == 1 else real,
                "fake_probability": fake, "real_probability": real}

    def analyze(self, text: Optional[str] = None, image_path: Optional[str] = None, video_path: Optional[str] = None,
                verbose: bool = True) -> Dict:
        say = print if verbose else (lambda *a, **k: None)
        say("\n" + "=" * 70 + "\nMISINFORMATION FORENSICS ANALYSIS\n" + "=" * 70)
        if not text and not image_path and not video_path:
            raise ValueError("Provide at least one of: text, image_path, or video_path")

        say("\n[Step 1] Text Analysis (RoBERTa Dual Heads)...")
        text_scores = {"ai_score": 0.0, "misinfo_score": 0.0}
        if text:
            text_scores = self.analyze_text(text)
            say(f"  • AI-Generated Score: {text_scores['ai_score']:.2%}")
            say(f"  • Misinfo/Propaganda Score: {text_scores['misinfo_score']:.2%}")
        else:
            say("  • Skipped (no text provided)")

        image_scores = {"deepfake_score": 0.0}
        consistency = {"clip_similarity": 0.0}
        vault = {"vault_discrepancy": 0.0, "matches": [], "vault_available": self.vault_loaded, "text_similarity": 0.0}
        if video_path:
            say("\n[Step 2] Video Forensics (Frame Sampling)...")
            v = self.analyze_video(video_path, text=text)
            image_scores["deepfake_score"] = v.get("deepfake_score", 0.0)
            consistency["clip_similarity"] = v.get("clip_similarity", 0.0)
            vault.update(vault_discrepancy=v.get("vault_discrepancy", 0.0), matches=v.get("vault_matches", []),
                         text_similarity=v.get("text_similarity", 0.0))
            say(f"  • Deepfake Probability (avg): {image_scores['deepfake_score']:.2%}")
            if text:
                say(f"  • CLIP Similarity (avg): {consistency['clip_similarity']:.4f}")
            say(f"  • Historical Discrepancy (max): {vault['vault_discrepancy']:.2%}")
        elif image_path:
            say("\n[Step 2] Visual Forensics (EfficientNet)...")
            image_scores = self.analyze_image(image_path)
            say(f"  • Deepfake Probability: {image_scores['deepfake_score']:.2%}")
            say("\n[Step 3] Image-Text Consistency (CLIP)...")
            shared_embed = None
            if text and self.dedup_clip_encode and self.vault_loaded:
                pil = self._to_pil_image(image_path)
                inputs = self.clip_processor(text=[text], images=pil, return_tensors="pt", padding=True).to(self.device)
                with torch.no_grad():
                    out = self.clip_model(**inputs)
                consistency = {"clip_similarity": self.engine.cosine_pairs(out.text_embeds, out.image_embeds).item()}
                shared_embed = out.image_embeds
                say(f"  • CLIP Similarity: {consistency['clip_similarity']:.4f}")
            elif text:
                consistency = self.analyze_consistency(text, image_path)
                say(f"  • CLIP Similarity: {consistency['clip_similarity']:.4f}")
            else:
                say("  • Skipped (no text provided)")
            say("\n[Step 4] Truth Vault Search (Guardian Database)...")
            vault = self._vault_results(shared_embed, [text], 5)[0] if shared_embed is not None else \
                self.search_vault(image_path, user_caption=text)
            if vault["vault_available"]:
                say(f"  • Historical Discrepancy: {vault['vault_discrepancy']:.2%}")
                if vault["matches"]:
                    say(f"  • Top Match: \"{vault['matches'][0]['title']}\"")
                    say(f"    Image Similarity: {vault['matches'][0]['similarity']:.1%}")
                    if vault.get("text_similarity", 0.0) > 0:
                        say(f"    Text Similarity: {vault['text_similarity']:.2%}")
            else:
                say("  • Vault not available")
        else:
            for step in ("[Step 2] Visual Forensics (EfficientNet)", "[Step 3] Image-Text Consistency (CLIP)",
                         "[Step 4] Truth Vault Search (Guardian Database)"):
                say(f"\n{step}...\n  • Skipped (no image/video provided)")

        all_scores = {**text_scores, **image_scores, **consistency,
                      "vault_discrepancy": vault["vault_discrepancy"], "text_similarity": vault.get("text_similarity", 0.0)}
        say("\n[Step 5] Verdict...")
        has_text, has_visual = bool(text), bool(image_path or video_path)
        verdict = self.fusion_verdict(all_scores) if (has_text and has_visual) else \
            self._fallback_verdict(all_scores, has_text, has_visual)
        all_scores.update(verdict)
        label = "FAKE" if verdict["verdict"] == 1 else "REAL"
        say(f"  {'🔴' if verdict['verdict'] == 1 else '🟢'} Final Verdict: {label}")
        say(f"  • Confidence: {verdict['confidence']:.1%}")

        say("\n[Step 6] Generating Forensic Summary...")
        explanation = self.generate_gemini_explanation(all_scores, vault["matches"])
        say("\n" + "=" * 70 + "\nFORENSIC SUMMARY\n" + "=" * 70 + f"\n{explanation}\n" + "=" * 70)
        return {"verdict": verdict["verdict"], "verdict_text": label, "confidence": verdict["confidence"],
                "scores": all_scores, "vault_matches": vault["matches"], "explanation": explanation}

    # ------------------------------------------------------------------ batched analyze (new surface)
    def analyze_batch(self, texts: Sequence[Optional[str]], images: Sequence[Optional[Union[str, Image.Image]]],
                      top_k: int = 5, encoder_batch: int = 64) -> List[Dict]:
        """analyze() for a batch of (text, image) samples (either may be None, not both): the producers run per
        modality in batches of `encoder_batch` (RoBERTa heads, EfficientNet, one batched CLIP forward per tower), then
        ONE pass of the hot path (cosine -> vault top-k -> fusion / fallback verdict) for the whole batch.  Per sample
        the result equals analyze(text, image_path) -- same keys, same values, up to the encoders' own batched-vs-single
        rounding (stock PyTorch, ~1e-6; encoder_batch=1 runs them exactly as analyze() does)."""
        n = len(texts)
        if len(images) != n:
            raise ValueError("texts and images must have the same length")
        head = np.zeros((n, 3), np.float32)
        mod = np.zeros(n, np.uint8)
        pils: List[Optional[Image.Image]] = [None] * n
        for i, (t, im) in enumerate(zip(texts, images)):
            if not t and im is None:
                raise ValueError("Provide at least one of: text, image_path, or video_path")
            if t:
                mod[i] |= 1
            if im is not None:
                pils[i] = self._to_pil_image(im)
                mod[i] |= 2
        with_text = [i for i in range(n) if mod[i] & 1]
        with_image = [i for i in range(n) if mod[i] & 2]
        if with_text:
            head[with_text, 0:2] = self.analyze_text_batch([texts[i] for i in with_text], encoder_batch)
        if with_image:
            head[with_image, 2] = self.analyze_image_batch([pils[i] for i in with_image], encoder_batch)
        t_emb = torch.zeros((n, 512), device=self.device)
        i_emb = torch.zeros((n, 512), device=self.device)
        vis = [i for i in range(n) if mod[i] & 2]
        both = [i for i in vis if mod[i] & 1]
        if vis:
            i_emb[vis] = self._clip_image_embed([pils[i] for i in vis]).float()
        if both:
            # no truncation, like analyze_consistency: a caption the reference cannot encode raises here too
            # (FusionTrainingDataset isolates such a sample and serves zeros for it, train_fusion_judge.py:97-99)
            t_emb[both] = self._clip_text_embed([texts[i] for i in both], truncation=False).float()
        out = self.score_batch(t_emb, i_emb, head, mod, top_k=top_k)
        x = out["scores"].cpu().numpy().astype(np.float64)
        probs, verdict, conf = out["probs"].cpu().numpy(), out["verdict"].cpu().numpy(), out["confidence"].cpu().numpy()
        vs, vr = out["vault_scores"].cpu().numpy(), out["vault_rows"].cpu().numpy()
        results, need = [], []
        for i in range(n):
            matches = self.vault.matches(vs[i], vr[i]) if (self.vault_loaded and mod[i] & 2) else []
            scores = {"ai_score": float(x[i, 0]), "misinfo_score": float(x[i, 1]), "deepfake_score": float(x[i, 2]),
                      "clip_similarity": float(x[i, 3]), "vault_discrepancy": float(x[i, 4]), "text_similarity": 0.0,
                      "verdict": int(verdict[i]), "confidence": float(conf[i]),
                      "fake_probability": float(probs[i, 1]), "real_probability": float(probs[i, 0])}
            if texts[i] and scores["vault_discrepancy"] != 0.0 and matches:
                need.append(i)
            results.append({"verdict": int(verdict[i]), "verdict_text": "FAKE" if verdict[i] == 1 else "REAL",
                            "confidence": float(conf[i]), "scores": scores, "vault_matches": matches, "explanation": ""})
        if need:   # caption vs matched headline, one batched text encode + one cosine launch
            emb = self._clip_text_embed([s for i in need for s in (texts[i], results[i]["vault_matches"][0]["title"])])
            for i, sim in zip(need, self.engine.cosine_pairs(emb[0::2], emb[1::2]).tolist()):
                results[i]["scores"]["text_similarity"] = float(sim)
        for r in results:
            r["explanation"] = self.generate_gemini_explanation(r["scores"], r["vault_matches"])
        return results

    def score_matrix(self, texts: Sequence[Optional[str]], images: Sequence[Optional[Union[str, Image.Image]]]) -> torch.Tensor:
        """(M,5) fusion-judge inputs [ai, misinfo, deepfake, clip_similarity, vault_discrepancy] for a dataset,
        computed ONCE -- what train_fusion_judge.FusionTrainingDataset.__getitem__ recomputes per sample per
        epoch through four analyze_* calls (train_fusion_judge.py:53-104)."""
        res = self.analyze_batch(texts, images)
        return torch.tensor([[r["scores"][k] for k in SCORE_ORDER] for r in res], dtype=torch.float32)

    def score_batch(self, text_embeds, image_embeds, head_scores, modality=None, top_k: int = 5):
        """Everything downstream of the encoders for a batch, without leaving the device:
        text_embeds/image_embeds (B,512), head_scores (B,3) = [ai, misinfo, deepfake],
        modality (B,) uint8 (bit0 text, bit1 visual; default both).  Per row the results equal
        the scalar path (analyze) on the same producer outputs.
        Returns dict of device tensors: clip_similarity, vault_discrepancy, vault_scores,
        vault_rows, probs (B,2) [real,fake], verdict, confidence."""
        return score_batch(self.engine, self.vault if self.vault_loaded else None, text_embeds, image_embeds,
                           head_scores, modality, top_k)
