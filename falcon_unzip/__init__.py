"""Drop-in alias of the reference's package name: `from falcon_unzip.phasing import main` (reference
src/py_scripts/fc_phasing.py:2) and pypeFLOW's by-name lookup of module-level task functions (falcon_unzip/unzip.py:304)
resolve to the B200 implementation in falcon_unzip_b200 without edits on the reference side.  Only the modules of the
hot path and of its widened neighbours exist; the workflow modules of the reference (unzip.py,
run_quiver.py, graphs_to_h_tigs.py, ...) are not replaced and keep importing from the reference's own tree."""
from falcon_unzip_b200 import __version__  # noqa: F401
