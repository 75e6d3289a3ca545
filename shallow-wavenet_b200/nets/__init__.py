from .cswnv_shift1 import CSWNV  # noqa: F401
from .dswnv import DSWNV, encode_mu_law, decode_mu_law  # noqa: F401
from ._base import initialize  # noqa: F401
