#!/usr/bin/env python
"""Entry point under the reference's script name (script/transfer_learning.R): ridge transfer of a fitted V to new
samples on the GPU.  See prmf_b200/transfer.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prmf_b200.transfer import main  # noqa: E402

if __name__ == "__main__":
    main()
