#!/usr/bin/env python
"""Entry script with the reference's name (script/nmf_pathway_rr.py): random restarts, one run per GPU at a time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from prmf_b200.restarts import main  # noqa: E402

if __name__ == "__main__":
    main()
