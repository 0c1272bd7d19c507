#!/usr/bin/env python
from falcon_unzip_b200.readmaps import main_phasing_readmap as main
import sys
if __name__ == "__main__":
    main(sys.argv)
