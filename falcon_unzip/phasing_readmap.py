"""falcon_unzip.phasing_readmap (reference falcon_unzip/phasing_readmap.py:8-75) -> falcon_unzip_b200.readmaps."""
from falcon_unzip_b200.readmaps import get_phasing_readmap  # noqa: F401
from falcon_unzip_b200.readmaps import parse_args_phasing_readmap as parse_args  # noqa: F401
from falcon_unzip_b200.readmaps import main_phasing_readmap as main  # noqa: F401
