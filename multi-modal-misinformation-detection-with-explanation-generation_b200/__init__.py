"""mmf_b200 -- B200-native (sm_100a) scoring hot path of the multi-modal misinformation
detector: CLIP caption/image cosine, Truth-Vault cosine top-k + discrepancy rule, and the
5-score fusion judge, behind the reference's Python surface.  See DESIGN.md."""
__version__ = "0.1.0"
