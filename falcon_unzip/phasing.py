"""falcon_unzip.phasing -> falcon_unzip_b200.phasing (same names, arguments and files as the reference module)."""
import sys

from falcon_unzip_b200 import phasing as _impl

sys.modules[__name__] = _impl
