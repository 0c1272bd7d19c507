#!/usr/bin/env python
from falcon_unzip_b200.phasing import main
import sys
if __name__ == "__main__":
    main(sys.argv)
