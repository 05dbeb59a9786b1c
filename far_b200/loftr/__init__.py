from .loftr import LoFTR  # noqa: F401
from .config import default_cfg, far_eval_cfg, upstream_loftr_cfg, full_cfg  # noqa: F401
from .transformer import (LoFTREncoderLayer, LocalFeatureTransformer, LinearAttention, CrossAttention, CrossBlock,  # noqa: F401
                          LocalFeatureTransformerRegressor, get_positional_encodings)
from .coarse_matching import CoarseMatching  # noqa: F401
from .fine_matching import FineMatching  # noqa: F401
from .fine_preprocess import FinePreprocess  # noqa: F401
from .position_encoding import PositionEncodingSine  # noqa: F401
