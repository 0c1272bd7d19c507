#!/usr/bin/env python
from falcon_unzip_b200.select_reads_from_bam import main
import sys
if __name__ == "__main__":
    main(sys.argv)
