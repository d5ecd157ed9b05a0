"""bwbble_b200 -- B200-native implementation of BWBBLE's read-mapping hot path."""
from .params import default_params, params_to_cli  # noqa: F401
from .index import BwtIndex, load_bwt, build_index  # noqa: F401
from .align import Aligner, AlignResult, align_reads, alns2sam  # noqa: F401
from ._lib import BwbError, Params  # noqa: F401
