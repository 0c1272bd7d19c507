"""falcon_unzip.ovlp_filter_with_phase -> falcon_unzip_b200.ovlp_filter_with_phase (same names, arguments and files as the reference module)."""
import sys

from falcon_unzip_b200 import ovlp_filter_with_phase as _impl

sys.modules[__name__] = _impl
