#!/usr/bin/env python3
"""Entry point with the reference launcher's name and flags
(./paf-baseband2power.py -a paf-baseband2power.conf -b <dir> -c 0 -d 0 -e 0 -f <file.dada>)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from paf_baseband2power_b200.launcher import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
