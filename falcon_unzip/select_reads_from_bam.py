"""falcon_unzip.select_reads_from_bam -> falcon_unzip_b200.select_reads_from_bam (same names, arguments and files as the reference module)."""
import sys

from falcon_unzip_b200 import select_reads_from_bam as _impl

sys.modules[__name__] = _impl
