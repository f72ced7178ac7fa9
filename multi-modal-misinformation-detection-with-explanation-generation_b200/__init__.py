"""mmf_b200 -- B200-native (sm_100a) scoring hot path of the multi-modal misinformation
detector: CLIP caption/image cosine, Truth-Vault cosine top-k + discrepancy rule, and the
5-score fusion judge, behind the reference's Python surface.  See DESIGN.md.

Importing the package does not touch CUDA; the C-ABI library is loaded (and required) as
soon as an Engine / MisinfoForensics / CLIPSimilarityEngine is constructed."""
__version__ = "0.1.0"

from ._lib import MMFError, library_path  # noqa: F401
from .engine import Engine, MATCH_THRESHOLD, VAULT_THRESHOLD  # noqa: F401
from .vault import ShardPlan, TruthVault, exchange_candidates, load_vault_file, read_vault_dict  # noqa: F401
from .pipeline import score_batch  # noqa: F401
from .forensics import MisinfoForensics, MultiModalMisinfoDetector  # noqa: F401
from .clip_similarity_engine import CLIPSimilarityEngine  # noqa: F401
from .similar_articles import search_similar_articles  # noqa: F401
from .fusion_data import FusionTrainingDataset  # noqa: F401
