"""Stand-in for x-transformers>=1.42.23: the five symbols the reference imports
(e2_tts.py:39-46). Semantics restated in SURVEY.md Appendix A.1-A.4. Test infrastructure only."""
from .x_transformers import Attention, FeedForward, RMSNorm, AdaptiveRMSNorm, RotaryEmbedding  # noqa
