"""falcon_unzip.rr_hctg_track -> falcon_unzip_b200.rr_hctg_track (same names, arguments and files as the reference module)."""
import sys

from falcon_unzip_b200 import rr_hctg_track as _impl

sys.modules[__name__] = _impl
