#!/usr/bin/env python
from falcon_unzip_b200.rr_hctg_track import main
import sys
if __name__ == "__main__":
    main(sys.argv)
