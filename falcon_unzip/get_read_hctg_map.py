"""falcon_unzip.get_read_hctg_map (reference falcon_unzip/get_read_hctg_map.py:12-103) -> falcon_unzip_b200.readmaps."""
from falcon_unzip_b200.readmaps import generate_read_to_hctg_map, get_read_hctg_map  # noqa: F401
from falcon_unzip_b200.readmaps import parse_args_get_read_hctg_map as parse_args  # noqa: F401
from falcon_unzip_b200.readmaps import main_get_read_hctg_map as main  # noqa: F401
