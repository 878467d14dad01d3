#!/usr/bin/env python
"""Entry script with the reference's name (setup.py installs script/prmf_runner.py on PATH)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from prmf_b200.prmf_runner import main  # noqa: E402

if __name__ == "__main__":
    main()
