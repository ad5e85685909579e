"""gdr_b200 — B200-native (sm_100a) fine-grained stage of GDR: cluster-restricted dense scoring +
top-k, and the docid logit masks of the constrained beam search, behind the reference's own
Python API (dense.py / main_models.py).  All compute runs in libgdr_b200.so; there is no CPU path."""
from . import _cabi  # noqa: F401
from .dense import DenseModel, DensePooler, compute_similarity  # noqa: F401
from .generation import DeviceTrie, TreeMask, build_logit_mask, flatten_trie, position_mask_, select_valid_embedding  # noqa: F401
from .main_models import EncoderModel, FineStage, Node, TreeBuilder, dec_2d, decode_token, encode_query, encode_single_newid  # noqa: F401
from .store import ClusterStore  # noqa: F401
from .pipeline import PipelinedRetriever, Ticket  # noqa: F401
from . import contrastive, expand, index_io  # noqa: F401

__version__ = "0.1.0"
